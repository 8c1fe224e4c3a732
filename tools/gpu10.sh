set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze8.log 2>&1; grep -E "^run 2|pairs/s|kernel warp|overflow reasons" gpurun_out/analyze8.log | tail -5
STAR_B200_LIB=$PWD/star_b200/lib/libstar_b200_v3.so STAR_B200_STITCH_CTAS_PER_SM=3 timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze8_v3.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze8_v3.log | tail -2
STAR_B200_LIB=$PWD/star_b200/lib/libstar_b200_v4.so STAR_B200_STITCH_CTAS_PER_SM=4 timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze8_v4.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze8_v4.log | tail -2
