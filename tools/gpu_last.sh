mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/last_pytest.log; cat gpurun_out/last_pytest.log
