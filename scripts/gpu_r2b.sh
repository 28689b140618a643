# round 2, call B: new 3-slot pipeline -- parity suite, then the device-resident leg with host trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -15 gpurun_out/r2b_pytest.log
DNZ_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 900 gpurun_out/r2b_bench.json; grep superbatch gpurun_out/r2b_bench.err | sed -n 300,306p
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench_e2e.json 2> gpurun_out/r2b_bench_e2e.err; tail -c 700 gpurun_out/r2b_bench_e2e.json
