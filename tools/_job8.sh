run() { echo "== attn $1 resid $2 swiglu $3"; FSB_PF_ATTN_MB=$1 FSB_PF_RESID_MB=$2 FSB_PF_SWIGLU_MB=$3 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-codec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('decode ms/frame', d['config']['ms_per_decode_frame'])"; }
run 0 0 0
run 0 0 48
run 0 0 24
run 0 32 0
run 48 0 0
run 48 32 48
run 64 32 48
FSB_PF_ATTN_MB=48 FSB_PF_RESID_MB=32 FSB_PF_SWIGLU_MB=48 timeout 300 python tools/trace_frame.py 2>&1 | tail -17
