timeout 300 python tools/probe/l2_resident_gemm.py 2>&1 | tail -6
timeout 300 python -m pytest tests/test_lm_gpu.py -m gpu -q --timeout 200 --tb=short -x -k frame_callback 2>&1 | tail -5
