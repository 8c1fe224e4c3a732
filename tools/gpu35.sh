mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/an35_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/an35_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/seed \1 stitch \2 total \3 ms_heavy \4/') | $(grep 'heavy kernel warp' gpurun_out/an35_$tag.log | tail -1 | cut -c1-110)"; }
run all STAR_B200_FLAT_STORE_ALL=1 STAR_B200_FLAT_DEBUG=1
grep "flat path" gpurun_out/an35_all.log | tail -1
run sd6 STAR_B200_SEED_CTAS_PER_SM=6
run sd8 STAR_B200_SEED_CTAS_PER_SM=8
