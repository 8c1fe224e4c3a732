mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/an34_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/an34_$tag.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/stitch \1 total \2 ms_heavy \3/') | $(grep 'heavy kernel warp' gpurun_out/an34_$tag.log | tail -1 | cut -c1-90)"; }
run s3 STAR_B200_FLAT_SETUP_CTAS_PER_SM=3
run s4 STAR_B200_FLAT_SETUP_CTAS_PER_SM=4
STAR_B200_FLAT_SETUP_CTAS_PER_SM=4 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "engine" 2>&1 | tail -12 > gpurun_out/t34.log; tail -2 gpurun_out/t34.log
