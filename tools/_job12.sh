timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 200 --tb=short -x -k "fused_residual" 2>&1 | tail -15
timeout 600 python tools/bench_res_unit.py 2>&1 | tail -9
timeout 600 python -m pytest tests/test_codec_gpu.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -6
