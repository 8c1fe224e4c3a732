mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_parity.py::test_engine_matches_oracle_tiny[std]" -m gpu -x -q > gpurun_out/sanitizer.log 2>&1
grep -E "Invalid|at 0x|stitch.cu|seed.cu|by thread|Address" gpurun_out/sanitizer.log | head -30
