set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python tools/analyze_chunk.py 262144 > gpurun_out/analyze1.log 2>&1; tail -12 gpurun_out/analyze1.log
STAR_B200_FAST_MAXW=512 STAR_B200_FAST_MAXP=512 STAR_B200_FAST_MAXTR=256 timeout 900 python tools/analyze_chunk.py 262144 > gpurun_out/analyze2.log 2>&1; tail -12 gpurun_out/analyze2.log
