# Round-2 GPU call 4:  gpurun --timeout 3000 -- 'bash tools/r02_gpu4.sh'
# seed search with per-item reconvergence of the groups, bigger window caps, split 48; full GPU test-suite; GRCh38-sized lines
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { tag=$1; shift; env "$@" timeout 900 python tools/analyze_chunk.py ${PAIRS:-1048576} ${MM:-0.005} ${RL:-100} > gpurun_out/g4_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/g4_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_window.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*/seed \1 tiers \2 stitch \3 total \4/')"; }
run base A=1
if ! grep -q '^run 2' gpurun_out/g4_base.log; then echo "base run failed"; tail -3 gpurun_out/g4_base.log; exit 1; fi
run lanes4 STAR_B200_SEED_GROUP_LANES=4
run lanes16 STAR_B200_SEED_GROUP_LANES=16
run ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
run sort16 STAR_B200_SEED_SORT_BITS=16
el chr21 sweeps done
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/g4_gpu_tests.log 2>&1; tail -4 gpurun_out/g4_gpu_tests.log
el gpu tests done
STAR_B200_DEBUG=1 timeout 2400 python bench.py --preset grch38 --steps 5 --warmup 3 > gpurun_out/g4_bench_grch38.json 2> gpurun_out/g4_bench_grch38.log
echo "grch38 rc=$?"; grep "overflow tier" gpurun_out/g4_bench_grch38.log | sort | uniq -c | head; head -c 700 gpurun_out/g4_bench_grch38.json; echo; python - <<'PY'
import json
d=json.loads(open("gpurun_out/g4_bench_grch38.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cli", d["cli_e2e"]["value"] if d.get("cli_e2e") else None, "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "stitch", d["roofline"]["stitch_kernel_ms"], "parity", d["parity_sample"]["diffs"], d["cli_e2e"]["parity_vs_reference"] if d.get("cli_e2e") else None)
print(d["cli_e2e"]["stage_times_from_Log_out"] if d.get("cli_e2e") else None)
PY
el grch38 bench done
export STAR_B200_BENCH_PRESET=grch38
run g38_base STAR_B200_DEBUG=1
grep "overflow tier" gpurun_out/g4_g38_base.log | tail -2
run g38_lanes4 STAR_B200_SEED_GROUP_LANES=4
run g38_oldlib STAR_B200_LIB=$PWD/star_b200/lib_ab/libstar_b200_pre_single_writer.so
PAIRS=262144 MM=0.05 RL=150 run g38_hard150 STAR_B200_DEBUG=1
grep "overflow tier" gpurun_out/g4_g38_hard150.log | tail -3; tail -8 gpurun_out/g4_g38_hard150.log | cut -c1-250
el grch38 sweeps done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g4_g38_launches.csv python bench.py --preset grch38 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g4_g38_launch_bench.log 2>&1
for k in seed_keyed_search_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/g4_g38_$k python bench.py --preset grch38 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g4_ncu_g38_$k.log 2>&1
  ncu -i gpurun_out/g4_g38_$k.ncu-rep --page raw --csv > gpurun_out/g4_g38_${k}_raw.csv 2>/dev/null
done
el ncu done
timeout 1500 python bench.py --preset grch38 --steps 3 --warmup 2 --read-len 150 --mm 0.05 --pairs 262144 --no-cli > gpurun_out/g4_bench_grch38_150.json 2> gpurun_out/g4_bench_grch38_150.log; head -c 300 gpurun_out/g4_bench_grch38_150.json; echo
el all done
