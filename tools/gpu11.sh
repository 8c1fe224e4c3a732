set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12
timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze9.log 2>&1; grep -E "^run 2|pairs/s|kernel warp|overflow reasons" gpurun_out/analyze9.log | tail -5
