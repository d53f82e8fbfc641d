timeout 300 python tools/trace_frame.py 2>&1 | tail -17
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-codec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('decode ms/frame', d['config']['ms_per_decode_frame'])"
timeout 600 python -m pytest tests -m gpu -q --timeout 200 --tb=short 2>&1 | tail -8
