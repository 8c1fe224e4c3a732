set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
for n in 262144 1048576; do timeout 900 python tools/analyze_chunk.py $n > gpurun_out/analyze6_$n.log 2>&1; grep -E "^run 2|pairs/s" gpurun_out/analyze6_$n.log; done
