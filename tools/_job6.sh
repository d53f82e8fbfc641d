timeout 300 python tools/trace_frame.py 2>&1 | tail -18
for v in "8 2" "16 2" "4 2" "8 4" "8 1"; do set -- $v; echo "== swiglu rb $1 resid rb $2"; FSB_SWIGLU_RB=$1 FSB_RESID_RB=$2 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-codec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('decode ms/frame', d['config']['ms_per_decode_frame'])"; done
timeout 600 python -m pytest tests -m gpu -q --timeout 200 --tb=short -x 2>&1 | tail -8
