mkdir -p gpurun_out/r02
FSB_CODEC_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/codec_launches.csv python tools/profile_codec.py 32 256 > gpurun_out/r02/pc.log 2>&1
python tools/summarize_launches.py gpurun_out/r02/codec_launches.csv 2>/dev/null | head -24
