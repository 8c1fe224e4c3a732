# Round-2 GPU call 6 (state of HEAD):  gpurun --timeout 2700 -- 'bash tools/r02_gpu6.sh'
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { tag=$1; shift; env "$@" timeout 900 python tools/analyze_chunk.py ${PAIRS:-1048576} ${MM:-0.005} ${RL:-100} > gpurun_out/g6_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/g6_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_window.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*/seed \1 tiers \2 stitch \3 total \4/')"; }
run base A=1
el chr21 done
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/g6_gpu_tests.log 2>&1; tail -3 gpurun_out/g6_gpu_tests.log
el gpu tests done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g6_smoke.log 2>&1; tail -1 gpurun_out/g6_smoke.log
STAR_B200_DEBUG=1 timeout 2400 python bench.py --preset grch38 --steps 10 --warmup 3 > gpurun_out/g6_bench_grch38.json 2> gpurun_out/g6_bench_grch38.log
echo "grch38 rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/g6_bench_grch38.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cli", d["cli_e2e"]["value"] if d.get("cli_e2e") else None, "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "stitch", d["roofline"]["stitch_kernel_ms"], "parity", d["parity_sample"]["diffs"], d["cli_e2e"]["parity_vs_reference"] if d.get("cli_e2e") else None)
print({k: d["cli_e2e"][k] for k in ("stage_times_from_Log_out", "wall_s", "startup_and_index_load_s", "mapping_pass_wall_s", "pairs_per_s_by_wall_minus_startup", "pairs_per_s_by_mapping_pass_wall")} if d.get("cli_e2e") else None)
PY
el grch38 bench done
timeout 900 python bench.py --impl reference --preset grch38 --steps 3 --warmup 1 > gpurun_out/g6_bench_reference.json 2> gpurun_out/g6_bench_reference.log; head -c 400 gpurun_out/g6_bench_reference.json; echo
el reference arm done
export STAR_B200_BENCH_PRESET=grch38
PAIRS=262144 MM=0.05 RL=150 run g38_hard150 STAR_B200_DEBUG=1
grep "overflow tier" gpurun_out/g6_g38_hard150.log | sort | uniq -c
el g38 150 done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g6_g38_launches.csv python bench.py --preset grch38 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g6_g38_launch_bench.log 2>&1
for k in seed_keyed_search_kernel flat_setup_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/g6_g38_$k python bench.py --preset grch38 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g6_ncu_g38_$k.log 2>&1
  ncu -i gpurun_out/g6_g38_$k.ncu-rep --page raw --csv > gpurun_out/g6_g38_${k}_raw.csv 2>/dev/null
done
el ncu done
timeout 1500 python bench.py --preset grch38 --steps 3 --warmup 2 --read-len 150 --mm 0.05 --pairs 262144 --no-cli > gpurun_out/g6_bench_grch38_150.json 2> gpurun_out/g6_bench_grch38_150.log; head -c 300 gpurun_out/g6_bench_grch38_150.json; echo
el all done
