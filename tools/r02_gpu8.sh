# Round-2 GPU call 8 (last):  gpurun --timeout 1500 -- 'bash tools/r02_gpu8.sh'
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/g8_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/g8_gpu_tests.log
el gpu tests done
timeout 1500 python bench.py --preset grch38 --steps 10 --warmup 3 > gpurun_out/g8_bench_grch38.json 2> gpurun_out/g8_bench_grch38.log
echo "grch38 rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/g8_bench_grch38.json").read().strip().split("\n")[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "cli", d["cli_e2e"]["value"] if d.get("cli_e2e") else None, "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "stitch", d["roofline"]["stitch_kernel_ms"], "parity", d["parity_sample"]["diffs"], d["cli_e2e"]["parity_vs_reference"] if d.get("cli_e2e") else None)
PY
el bench done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:seed_keyed_search_kernel -s 1 -c 1 -o gpurun_out/g8_g38_seed python bench.py --preset grch38 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g8_ncu_g38_seed.log 2>&1
ncu -i gpurun_out/g8_g38_seed.ncu-rep --page raw --csv > gpurun_out/g8_g38_seed_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/g8_g38_seed_raw.csv"))); hdr=rows[0]; vals=rows[2]
for k in ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio"]: print(k, vals[hdr.index(k)])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g8_g38_launches.csv python bench.py --preset grch38 --steps 2 --warmup 1 --no-cli --no-cpu > gpurun_out/g8_launch_bench.log 2>&1
el launch list done
timeout 600 python tools/variants.py grch38 1048576 -- base STAR_B200_BIN_FILTER=0 STAR_B200_L2_FETCH_BYTES=64 STAR_B200_SORTED_LOOKUP_MIN=48 STAR_B200_SORTED_LOOKUP_MIN=4 STAR_B200_FLAT_SETUP_CTAS_PER_SM=4 STAR_B200_FLAT_DFS_CTAS_PER_SM=5 STAR_B200_SEED_KEYED_CTAS_PER_SM=12 STAR_B200_HEAVY_SPLIT=40 base > gpurun_out/g8_variants.jsonl 2> gpurun_out/g8_variants.log; cat gpurun_out/g8_variants.jsonl
el variants done
el all done
