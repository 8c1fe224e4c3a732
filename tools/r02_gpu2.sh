# Round-2 GPU call 2:  gpurun --timeout 3000 -- 'bash tools/r02_gpu2.sh'
# keyed seed stage at 1 M pairs (chr21-sized, then GRCh38-sized): timing, sweeps, ncu captures, both bench lines
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { tag=$1; shift; env "$@" timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/g2_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/g2_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*ms_heavy.: ([0-9.]+).*/seed \1 stitch \2 total \3 ms_heavy \4/')"; }
run base A=1
if ! grep -q '^run 2' gpurun_out/g2_base.log; then
  echo "base run failed: memcheck"; tail -3 gpurun_out/g2_base.log
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/analyze_chunk.py 1048576 > gpurun_out/g2_memcheck.log 2>&1
  grep -E "Invalid|at 0x|by thread|Address" gpurun_out/g2_memcheck.log | head -30
  exit 1
fi
el base done
run sort0 STAR_B200_SEED_SORT_BITS=0
run sort28 STAR_B200_SEED_SORT_BITS=28
run lanes4 STAR_B200_SEED_GROUP_LANES=4
run lanes16 STAR_B200_SEED_GROUP_LANES=16
run ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
run lanes4ctas12 STAR_B200_SEED_GROUP_LANES=4 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
for s in 20 40; do run split$s STAR_B200_HEAVY_SPLIT=$s; done
run dfs6 STAR_B200_FLAT_DFS_CTAS_PER_SM=6
run setup4 STAR_B200_FLAT_SETUP_CTAS_PER_SM=4
run storeall0 STAR_B200_FLAT_STORE_ALL=0
el chr21 sweeps done
# launch list + full captures of the hot kernels (262144 pairs keep the replays short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g2_launches.csv python bench.py --preset chr21 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g2_launch_bench.log 2>&1
for k in flat_dfs_warp_kernel flat_setup_kernel flat_record_warp_kernel seed_keyed_search_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/g2_$k python bench.py --preset chr21 --pairs 262144 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g2_ncu_$k.log 2>&1
  ncu -i gpurun_out/g2_$k.ncu-rep --page raw --csv > gpurun_out/g2_${k}_raw.csv 2>/dev/null
done
el ncu chr21 done
timeout 900 python bench.py --preset chr21 --steps 5 --warmup 3 > gpurun_out/g2_bench_chr21.json 2> gpurun_out/g2_bench_chr21.log; tail -c 1800 gpurun_out/g2_bench_chr21.json; echo
el chr21 bench done
# GRCh38-sized: index build + full bench line, then sweeps of the seed stage and one capture at real size
STAR_B200_DEBUG=1 STAR_B200_SA_DEBUG=1 timeout 2400 python bench.py --preset grch38 --steps 5 --warmup 3 > gpurun_out/g2_bench_grch38.json 2> gpurun_out/g2_bench_grch38.log
echo "grch38 rc=$?"; tail -4 gpurun_out/g2_bench_grch38.log | cut -c1-300; tail -c 3500 gpurun_out/g2_bench_grch38.json; echo
cp /tmp/star_b200_bench/grch38/gen_Log.out gpurun_out/g2_grch38_gen_Log.out 2>/dev/null
el grch38 bench done
export STAR_B200_BENCH_PRESET=grch38
run g38_base A=1
run g38_sort0 STAR_B200_SEED_SORT_BITS=0
run g38_sort28 STAR_B200_SEED_SORT_BITS=28
run g38_lanes4 STAR_B200_SEED_GROUP_LANES=4
run g38_ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
el grch38 sweeps done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_keyed_search_kernel -s 1 -c 1 -o gpurun_out/g2_g38_seed_keyed_search_kernel python bench.py --preset grch38 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g2_ncu_g38_seed.log 2>&1
ncu -i gpurun_out/g2_g38_seed_keyed_search_kernel.ncu-rep --page raw --csv > gpurun_out/g2_g38_seed_keyed_search_kernel_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g2_g38_launches.csv python bench.py --preset grch38 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g2_g38_launch_bench.log 2>&1
el all done
