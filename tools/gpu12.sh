set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze10.log 2>&1; grep -E "^run 2|pairs/s|kernel warp" gpurun_out/analyze10.log | tail -4
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stitch_heavy_kernel -c 1 -o gpurun_out/prof_heavy -f python tools/analyze_chunk.py 262144 > gpurun_out/ncu_heavy.log 2>&1; tail -2 gpurun_out/ncu_heavy.log
