# Round-2 GPU call 7 (2 GPUs):  gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_gpu7.sh'
# bench under torchrun at N=2 (GRCh38-sized index) + the product multi-GPU path (python -m star_b200.dist) at N=1 and N=2, plain and 2-pass
mkdir -p gpurun_out
export STAR_B200_BENCH_DIR=/tmp/star_b200_bench
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi -L | head -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g7_bench_n2.json 2> gpurun_out/g7_bench_n2.log
echo "bench n2 rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/g7_bench_n2.json").read().strip().split("\n")[-1])
    print("N=2 value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity_sample"]["diffs"], "workload", d["config"]["workload"][:40], d["config"].get("fallback"))
except Exception as e: print("no line", e)
PY
tail -3 gpurun_out/g7_bench_n2.log | cut -c1-200
el bench n2 done
timeout 600 python tools/product_scale.py prepare --preset grch38 --pairs 8000000 > gpurun_out/g7_prod_prepare.json 2> gpurun_out/g7_prod_prepare.log; tail -c 300 gpurun_out/g7_prod_prepare.json
el prepare done
for n in 1 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n tools/product_scale.py run --preset grch38 --mode map > gpurun_out/g7_prod_map_n$n.json 2> gpurun_out/g7_prod_map_n$n.log; tail -c 700 gpurun_out/g7_prod_map_n$n.json; echo
done
el product map done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/product_scale.py run --preset grch38 --mode twopass > gpurun_out/g7_prod_twopass_n2.json 2> gpurun_out/g7_prod_twopass_n2.log; tail -c 900 gpurun_out/g7_prod_twopass_n2.json; echo; tail -3 gpurun_out/g7_prod_twopass_n2.log | cut -c1-200
el all done
