mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_pytest.log; tail -2 gpurun_out/final_pytest.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2500 gpurun_out/bench_final.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; tail -c 900 gpurun_out/bench_final_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_search_kernel -c 1 -o gpurun_out/r01_seed_full -f python tools/analyze_chunk.py 1048576 > gpurun_out/ncu_seed_final.log 2>&1; tail -1 gpurun_out/ncu_seed_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_dfs_warp_kernel -c 1 -o gpurun_out/r01_dfs_full -f python tools/analyze_chunk.py 1048576 > gpurun_out/ncu_dfs_final.log 2>&1; tail -1 gpurun_out/ncu_dfs_final.log
