mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity_large.py > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 700 gpurun_out/r2e_bench.json; echo
DNZ_TRACE=1 timeout 600 python bench.py --workload cfg1 --steps 3 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/r2e_cfg1.json 2> gpurun_out/r2e_cfg1.err; tail -c 300 gpurun_out/r2e_cfg1.json; grep superbatch gpurun_out/r2e_cfg1.err | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_aggregate -s 2 -c 1 -o gpurun_out/prof_agg_r2e -f python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2e_ncu.log 2>&1
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2e_cfg5.json 2> gpurun_out/r2e_cfg5.err; tail -c 400 gpurun_out/r2e_cfg5.json; tail -2 gpurun_out/r2e_cfg5.err
