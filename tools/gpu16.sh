mkdir -p gpurun_out
STAR_B200_LIB=$PWD/star_b200/lib/libstar_b200_v3.so STAR_B200_STITCH_CTAS_PER_SM=3 timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze13_v3.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze13_v3.log | tail -2 | cut -c1-400
STAR_B200_LIB=$PWD/star_b200/lib/libstar_b200_v4.so STAR_B200_STITCH_CTAS_PER_SM=4 timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze13_v4.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze13_v4.log | tail -2 | cut -c1-400
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stitch_heavy_kernel -c 1 -o gpurun_out/prof_heavy2 -f python tools/analyze_chunk.py 262144 > gpurun_out/ncu_heavy2.log 2>&1; tail -1 gpurun_out/ncu_heavy2.log
