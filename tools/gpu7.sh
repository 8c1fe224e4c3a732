set -x
mkdir -p gpurun_out
which ncu; ncu --version | tail -1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:stitch_kernel -c 1 -o gpurun_out/prof_stitch -f python tools/analyze_chunk.py 131072 > gpurun_out/ncu_stitch.log 2>&1; tail -3 gpurun_out/ncu_stitch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:seed_search_kernel -c 1 -o gpurun_out/prof_seed -f python tools/analyze_chunk.py 1048576 > gpurun_out/ncu_seed.log 2>&1; tail -3 gpurun_out/ncu_seed.log
ls -la gpurun_out/
