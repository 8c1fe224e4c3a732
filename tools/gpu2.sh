set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 1200 python bench.py --steps 3 --warmup 1 --pairs 262144 --ref-pairs 1000000 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.log
tail -5 gpurun_out/bench_first.log; cat gpurun_out/bench_first.json
