set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze11.log 2>&1; grep -E "^run 2|pairs/s|kernel warp|memo|E-phase" gpurun_out/analyze11.log | tail -6
