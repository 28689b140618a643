mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_large.py -x -q --durations=6 > gpurun_out/r2d_pytest_large.log 2>&1; tail -14 gpurun_out/r2d_pytest_large.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 1500 gpurun_out/r2d_bench.json; grep -E "oracle|equal|Error|error" gpurun_out/r2d_bench.err | tail
ncu --set full --clock-control none --import-source on -k regex:k_aggregate -s 2 -c 1 -o gpurun_out/prof_agg_r2d -f python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
for wl in cfg3 cfg5 cfg1 cfg4; do
  timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2d_$wl.json 2> gpurun_out/r2d_$wl.err; tail -c 400 gpurun_out/r2d_$wl.json; echo
done
