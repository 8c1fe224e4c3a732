mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "engine" 2>&1 | tail -12 > gpurun_out/t29.log; tail -2 gpurun_out/t29.log
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/an29_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/an29_$tag.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/stitch \1 total \2 ms_heavy \3/') | $(grep 'heavy kernel warp' gpurun_out/an29_$tag.log | tail -1 | cut -c1-90)"; }
run w4 STAR_B200_FLAT_DFS_CTAS_PER_SM=4
run w4s12 STAR_B200_FLAT_DFS_CTAS_PER_SM=4 STAR_B200_HEAVY_SPLIT=12
run w6s12 STAR_B200_FLAT_DFS_CTAS_PER_SM=6 STAR_B200_HEAVY_SPLIT=12
run w4s12na32 STAR_B200_FLAT_DFS_CTAS_PER_SM=4 STAR_B200_HEAVY_SPLIT=12 STAR_B200_HEAVY_NA=32
run w4s12na16 STAR_B200_FLAT_DFS_CTAS_PER_SM=4 STAR_B200_HEAVY_SPLIT=12 STAR_B200_HEAVY_NA=16
