timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 200 --tb=short -x -k "fused_residual" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_codec_gpu.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -8
for v in 1 0; do echo "== fused $v"; FSB_FUSED_RESUNIT=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('codec ms', d['config']['codec_ms'], 'value', d['value'])"; done
