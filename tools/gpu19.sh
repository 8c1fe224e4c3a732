mkdir -p gpurun_out
timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze14.log 2>&1; grep -E "^run 2|pairs/s|kernel warp|E-phase" gpurun_out/analyze14.log | tail -5
