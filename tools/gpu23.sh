mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "engine" 2>&1 | tail -15 > gpurun_out/t23.log; tail -8 gpurun_out/t23.log
STAR_B200_FLAT_DEBUG=1 timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze23.log 2>&1; grep -E "^run 2|pairs/s|flat path|heavy kernel|E-phase" gpurun_out/analyze23.log | tail -6 | cut -c1-600
