mkdir -p gpurun_out
for na in 32 128 256 512 2048; do
  STAR_B200_HEAVY_NA=$na timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/sweep_na$na.log 2>&1
  echo "NA=$na $(grep -E '^run 2' gpurun_out/sweep_na$na.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*slow_path_reads.: ([0-9]+), .heavy_reads.: ([0-9]+), .ms_heavy.: ([0-9.]+).*/stitch \1 total \2 slow \3 heavy \4 ms_heavy \5/')"
done
for est in 256 4096 16384; do
  STAR_B200_HEAVY_EST=$est timeout 600 python tools/analyze_chunk.py 1048576 > gpurun_out/sweep_est$est.log 2>&1
  echo "EST=$est $(grep -E '^run 2' gpurun_out/sweep_est$est.log | sed -E 's/.*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*slow_path_reads.: ([0-9]+), .heavy_reads.: ([0-9]+), .ms_heavy.: ([0-9.]+).*/stitch \1 total \2 slow \3 heavy \4 ms_heavy \5/')"
done
