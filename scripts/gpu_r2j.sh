mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity_large.py > gpurun_out/r2j_pytest.log 2>&1; tail -6 gpurun_out/r2j_pytest.log
