mkdir -p gpurun_out
for sp in 8 10 12; do
STAR_B200_HEAVY_SPLIT=$sp timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze16_$sp.log 2>&1; echo "SPLIT=$sp $(grep -E '^run 2' gpurun_out/analyze16_$sp.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*heavy_reads.: ([0-9]+), .ms_heavy.: ([0-9.]+).*/stitch \1 total \2 heavy \3 ms_heavy \4/') $(grep 'heavy kernel warp' gpurun_out/analyze16_$sp.log | tail -1 | sed 's/.*tasks/tasks/')"
done
