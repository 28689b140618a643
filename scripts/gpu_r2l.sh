mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ungrouped.py -q > gpurun_out/r2l_pytest_u.log 2>&1; tail -25 gpurun_out/r2l_pytest_u.log
