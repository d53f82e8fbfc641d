mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_lm_gpu.py -m gpu -q --timeout 200 -x -k "frame_callback" 2>&1 | tail -5
for k in 0 2 4 8 16; do
  echo "== L2 prefetch $k"
  FSB_L2_PREFETCH=$k timeout 300 python tools/trace_frame.py 2>&1 | head -14 > gpurun_out/r02/trace_l2pf_$k.md; head -1 gpurun_out/r02/trace_l2pf_$k.md; tail -1 gpurun_out/r02/trace_l2pf_$k.md
  FSB_L2_PREFETCH=$k timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-codec 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('decode ms/frame', d['config']['ms_per_decode_frame'])"
done
