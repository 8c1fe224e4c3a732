# Round-2 GPU call 5:  gpurun --timeout 2400 -- 'bash tools/r02_gpu5.sh'
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { tag=$1; shift; env "$@" timeout 900 python tools/analyze_chunk.py ${PAIRS:-1048576} ${MM:-0.005} ${RL:-100} > gpurun_out/g5_$tag.log 2>&1; echo "$tag $(grep -E '^run 2' gpurun_out/g5_$tag.log | sed -E 's/.*ms_seed.: ([0-9.]+).*ms_window.: ([0-9.]+).*ms_stitch.: ([0-9.]+).*ms_total.: ([0-9.]+).*/seed \1 tiers \2 stitch \3 total \4/')"; }
run base A=1
run ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
run lanes4 STAR_B200_SEED_GROUP_LANES=4
el chr21 done
STAR_B200_DEBUG=1 timeout 2400 python bench.py --preset grch38 --steps 5 --warmup 3 > gpurun_out/g5_bench_grch38.json 2> gpurun_out/g5_bench_grch38.log
echo "grch38 rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/g5_bench_grch38.json").read().strip().split("\n")[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cli", d["cli_e2e"]["value"] if d.get("cli_e2e") else None, "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "stitch", d["roofline"]["stitch_kernel_ms"], "parity", d["parity_sample"]["diffs"], d["cli_e2e"]["parity_vs_reference"] if d.get("cli_e2e") else None)
print(d["cli_e2e"]["stage_times_from_Log_out"] if d.get("cli_e2e") else None, d["cli_e2e"]["wall_s"], d["cli_e2e"]["startup_and_index_load_s"])
PY
el grch38 bench done
export STAR_B200_BENCH_PRESET=grch38
run g38_base A=1
run g38_ctas12 STAR_B200_SEED_KEYED_CTAS_PER_SM=12
PAIRS=262144 MM=0.05 RL=150 run g38_hard150 STAR_B200_DEBUG=1
grep "overflow tier" gpurun_out/g5_g38_hard150.log | tail -2
PAIRS=262144 MM=0.05 RL=150 run g38_hard150_nosort STAR_B200_SORTED_LOOKUP_MIN=100000
el g38 runs done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/g5_g38_150_launches.csv python tools/analyze_chunk.py 262144 0.05 150 > gpurun_out/g5_g38_150_launch.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open("gpurun_out/g5_g38_150_launches.csv")) if len(r)>5]
for i,r in enumerate(rows):
    if "Kernel Name" in r: hdr=r; start=i; break
ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[start+2:]:
    try: name=r[ki]; v=float(r[vi].replace(",",""))
    except: continue
    name=re.sub(r"\(.*","",name)[:60]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print("  %-60s n=%3d  %.2f ms/launch total %.1f ms"%(k,n,v/n/1e6,v/1e6))
PY
el launch list done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_keyed_search_kernel -s 1 -c 1 -o gpurun_out/g5_g38_seed python bench.py --preset grch38 --steps 1 --warmup 1 --no-cli --no-cpu > gpurun_out/g5_ncu_g38_seed.log 2>&1
ncu -i gpurun_out/g5_g38_seed.ncu-rep --page raw --csv > gpurun_out/g5_g38_seed_raw.csv 2>/dev/null
el all done
