mkdir -p gpurun_out
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 --ref-pairs 1000000 --ref-repeat 16 > gpurun_out/ref16.json 2> gpurun_out/ref16.log; cat gpurun_out/ref16.json | cut -c1-300; tail -2 gpurun_out/ref16.log
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 --ref-pairs 1000000 --ref-repeat 4 > gpurun_out/ref4.json 2> gpurun_out/ref4.log; cat gpurun_out/ref4.json | cut -c1-300; tail -2 gpurun_out/ref4.log
