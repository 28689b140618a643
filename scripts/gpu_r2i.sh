mkdir -p gpurun_out
DNZ_TRACE=1 timeout 900 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2i_cfg5.json 2> gpurun_out/r2i_cfg5.err; tail -c 500 gpurun_out/r2i_cfg5.json; grep -c superbatch gpurun_out/r2i_cfg5.err; grep "alloc" gpurun_out/r2i_cfg5.err | tail -40; tail -3 gpurun_out/r2i_cfg5.err
