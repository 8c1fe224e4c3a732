mkdir -p gpurun_out
for mw in 256 512; do
STAR_B200_HEAVY_MAXW=$mw timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/analyze15_$mw.log 2>&1; echo "MAXW=$mw $(grep -E '^run 2' gpurun_out/analyze15_$mw.log | sed -E 's/.*ms_window.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*slow_path_reads.: ([0-9]+), .heavy_reads.: ([0-9]+), .ms_heavy.: ([0-9.]+).*/tiers \1 stitch \2 total \3 slow \4 heavy \5 ms_heavy \6/')"
done
