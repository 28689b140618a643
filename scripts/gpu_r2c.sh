mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -15 gpurun_out/r2c_pytest.log
DNZ_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 900 gpurun_out/r2c_bench.json; grep superbatch gpurun_out/r2c_bench.err | sed -n 329,336p
