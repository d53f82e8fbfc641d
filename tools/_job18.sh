mkdir -p gpurun_out/r02f
timeout 900 python bench.py > gpurun_out/r02f/bench.json 2> gpurun_out/r02f/bench.err
for w in single roundtrip stream voice-clone serve; do timeout 400 python bench.py --workload $w > gpurun_out/r02f/bench_$w.json 2> gpurun_out/r02f/bench_$w.err; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f/frame_launches.csv python tools/profile_frame.py --frames 2 > gpurun_out/r02f/pf.log 2>&1
timeout 300 python tools/trace_frame.py > gpurun_out/r02f/trace_frame.md 2>&1
timeout 300 python tools/bench_res_unit.py > gpurun_out/r02f/res_unit.md 2>&1
timeout 300 python tools/probe/l2_resident_gemm.py > gpurun_out/r02f/l2_resident.md 2>&1
cut -c1-400 gpurun_out/r02f/bench.json; tail -3 gpurun_out/r02f/bench.err
