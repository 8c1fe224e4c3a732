mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.log; tail -3 gpurun_out/bench_r01.log; cat gpurun_out/bench_r01.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.log; cat gpurun_out/bench_r01_ref.json | cut -c1-400
