set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python tools/analyze_chunk.py 262144 > gpurun_out/analyze3.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze3.log
STAR_B200_LIB=$PWD/star_b200/lib/libstar_b200_v3.so STAR_B200_STITCH_CTAS_PER_SM=3 timeout 900 python tools/analyze_chunk.py 262144 > gpurun_out/analyze3_v3.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze3_v3.log
STAR_B200_LIB=$PWD/star_b200/lib/libstar_b200_v4.so STAR_B200_STITCH_CTAS_PER_SM=4 timeout 900 python tools/analyze_chunk.py 262144 > gpurun_out/analyze3_v4.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze3_v4.log
