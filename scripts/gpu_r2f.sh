mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_exchange.py -x -q > gpurun_out/r2f_pytest_x.log 2>&1; tail -12 gpurun_out/r2f_pytest_x.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.json; echo
timeout 200 python bench.py --workload cfg1 --steps 5 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/r2f_cfg1.json 2> gpurun_out/r2f_cfg1.err; tail -c 300 gpurun_out/r2f_cfg1.json
