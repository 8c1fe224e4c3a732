mkdir -p gpurun_out
STAR_B200_HEAVY_SPLIT=12 STAR_B200_FLAT_DFS_MODE=1 STAR_B200_FLAT_DFS_CTAS_PER_SM=4 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:flat_dfs_warp_kernel -c 1 -o gpurun_out/prof_flat_warp3 -f python tools/analyze_chunk.py 262144 > gpurun_out/ncu_flat_warp3.log 2>&1; tail -1 gpurun_out/ncu_flat_warp3.log
