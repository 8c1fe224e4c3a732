# Round-2 GPU call 3:  gpurun --timeout 3000 -- 'bash tools/r02_gpu3.sh'
# after: word-packed seed search, flat overflow tier, bulk skip + SA prefetch in the window assignment, multi-block scan
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { tag=$1; shift; env "$@" timeout 900 python tools/analyze_chunk.py 1048576 > gpurun_out/g3_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/g3_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_window.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*/seed \1 tiers \2 stitch \3 total \4/')"; }
run base A=1
if ! grep -q '^run 2' gpurun_out/g3_base.log; then echo "base run failed"; tail -3 gpurun_out/g3_base.log; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/analyze_chunk.py 262144 > gpurun_out/g3_memcheck.log 2>&1; grep -E "Invalid|at 0x|by thread|Address" gpurun_out/g3_memcheck.log | head -30; exit 1; fi
el base done
run sort16 STAR_B200_SEED_SORT_BITS=16
run lanes4 STAR_B200_SEED_GROUP_LANES=4
run ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
for s in 40 44 48 52; do run split$s STAR_B200_HEAVY_SPLIT=$s; done
run split44dfs3 STAR_B200_HEAVY_SPLIT=44 STAR_B200_FLAT_DFS_CTAS_PER_SM=3
el chr21 sweeps done
# GPU test-suite (everything changed underneath it)
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/g3_gpu_tests.log 2>&1; tail -4 gpurun_out/g3_gpu_tests.log
el gpu tests done
# GRCh38-sized
STAR_B200_DEBUG=1 timeout 2400 python bench.py --preset grch38 --steps 5 --warmup 3 > gpurun_out/g3_bench_grch38.json 2> gpurun_out/g3_bench_grch38.log
echo "grch38 rc=$?"; tail -3 gpurun_out/g3_bench_grch38.log | cut -c1-300; tail -c 3800 gpurun_out/g3_bench_grch38.json; echo
el grch38 bench done
export STAR_B200_BENCH_PRESET=grch38
run g38_base A=1
run g38_split44 STAR_B200_HEAVY_SPLIT=44
run g38_setup4 STAR_B200_FLAT_SETUP_CTAS_PER_SM=4
run g38_setup2 STAR_B200_FLAT_SETUP_CTAS_PER_SM=2
run g38_ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
run g38_lanes4 STAR_B200_SEED_GROUP_LANES=4
el grch38 sweeps done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g3_g38_launches.csv python bench.py --preset grch38 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g3_g38_launch_bench.log 2>&1
for k in seed_keyed_search_kernel flat_setup_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/g3_g38_$k python bench.py --preset grch38 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g3_ncu_g38_$k.log 2>&1
  ncu -i gpurun_out/g3_g38_$k.ncu-rep --page raw --csv > gpurun_out/g3_g38_${k}_raw.csv 2>/dev/null
done
el ncu done
# config 3 shape at GRCh38 size: 2x150 at 5 % (overflow telemetry in the line)
timeout 1500 python bench.py --preset grch38 --steps 3 --warmup 2 --read-len 150 --mm 0.05 --pairs 262144 --no-cli > gpurun_out/g3_bench_grch38_150.json 2> gpurun_out/g3_bench_grch38_150.log; tail -c 2500 gpurun_out/g3_bench_grch38_150.json; echo
el all done
