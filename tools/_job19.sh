mkdir -p gpurun_out/r02f
FSB_CODEC_GRAPH=0 timeout 900 ncu --set full --clock-control none -k regex:"res_unit_kernel|gemm_tc_kernel|attn_tile_kernel" -s 61 -c 61 -o /tmp/codec -f python tools/profile_codec.py 32 256 > gpurun_out/r02f/ncu2.log 2>&1
ncu -i /tmp/codec.ncu-rep --page raw --csv > gpurun_out/r02f/codec_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/r02f/codec_ncu_raw.csv; tail -n 2 gpurun_out/r02f/ncu2.log
