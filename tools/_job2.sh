mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_lm_gpu.py tests/test_codec_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout 200 -x -k "frame_callback or streaming or full_s2pro or chunked or bulk" 2>&1 | tail -40 > gpurun_out/r02/tests_a.log
timeout 600 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -15 > gpurun_out/r02/tests_all.log
timeout 600 python bench.py > gpurun_out/r02/bench.json 2> gpurun_out/r02/bench.err
for w in single roundtrip stream voice-clone serve; do timeout 400 python bench.py --workload $w > gpurun_out/r02/bench_$w.json 2> gpurun_out/r02/bench_$w.err; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/frame_launches.csv python tools/profile_frame.py --frames 2 > gpurun_out/r02/pf.log 2>&1

timeout 300 python tools/trace_frame.py --every 40 > gpurun_out/r02/trace_frame.md 2>&1; head -20 gpurun_out/r02/trace_frame.md
tail -5 gpurun_out/r02/tests_a.log; tail -3 gpurun_out/r02/tests_all.log; cut -c1-600 gpurun_out/r02/bench.json
