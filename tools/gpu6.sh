set -x
mkdir -p gpurun_out
ANALYZE_LIGHT=1 timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze5.log 2>&1; grep -E "^run 2|pairs/s|light" gpurun_out/analyze5.log
