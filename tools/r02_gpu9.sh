# Round-2 GPU call 9 (last minutes of the budget):  gpurun --timeout 420 -- 'bash tools/r02_gpu9.sh'
# bench line at the final defaults (64-byte L2 fills, no bin filter) + the command line at three host-thread counts with per-stage times
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python bench.py --preset grch38 --steps 5 --warmup 3 --no-cli --no-cpu > gpurun_out/g9_bench_grch38.json 2> gpurun_out/g9_bench_grch38.log
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/g9_bench_grch38.json").read().strip().split("\n")[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "traffic", d["roofline"]["traffic"], "parity", d["parity_sample"]["diffs"])
PY
el bench done
timeout 120 python tools/product_scale.py prepare --preset grch38 --pairs 4000000 > gpurun_out/g9_prepare.json 2> gpurun_out/g9_prepare.log
el fastq done
W=/tmp/star_b200_bench/grch38
for T in 64 32 112; do
  mkdir -p $W/cli_t$T
  STAR_B200_READER_DEBUG=1 timeout 60 star_b200/bin/STAR --genomeDir $W/idx --readFilesIn $W/prod_1.fq $W/prod_2.fq --runThreadN $T --outSAMtype SAM --outFileNamePrefix $W/cli_t$T/ > /dev/null 2> gpurun_out/g9_cli_t$T.err
  echo "T=$T rc=$?"; grep "star-b200:" $W/cli_t$T/Log.out; grep "^reader:" gpurun_out/g9_cli_t$T.err | tail -2
  grep "star-b200:" $W/cli_t$T/Log.out > gpurun_out/g9_cli_t$T.txt
  rm -f $W/cli_t$T/Aligned.out.sam
  el cli $T done
done
el all done
