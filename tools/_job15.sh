timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -5
timeout 300 python tools/profile_codec.py 32 256 3 2>&1 | tail -2
