mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/an24_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/an24_$tag.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/stitch \1 total \2 ms_heavy \3/') | $(grep 'heavy kernel warp' gpurun_out/an24_$tag.log | tail -1 | cut -c1-90) | $(grep 'E-phase' gpurun_out/an24_$tag.log | tail -1)"; }
run c3 STAR_B200_FLAT_DFS_CTAS_PER_SM=3
run c4 STAR_B200_FLAT_DFS_CTAS_PER_SM=4
run ls2 STAR_B200_FLAT_LANE_SCRATCH=1
run ls4 STAR_B200_FLAT_LANE_SCRATCH=1 STAR_B200_FLAT_DFS_CTAS_PER_SM=4
run f32 STAR_B200_FLAT_FETCH_MIN=32
run f16 STAR_B200_FLAT_FETCH_MIN=16
run f32c4 STAR_B200_FLAT_FETCH_MIN=32 STAR_B200_FLAT_DFS_CTAS_PER_SM=4
run f32c4ls STAR_B200_FLAT_FETCH_MIN=32 STAR_B200_FLAT_DFS_CTAS_PER_SM=4 STAR_B200_FLAT_LANE_SCRATCH=1
