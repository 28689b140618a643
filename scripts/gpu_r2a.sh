# round 2, call A: texture-path microbenchmark + host-timeline trace of the device-resident leg (baseline = round-1 build)
mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/r2a_host.txt
./profiles/microbench/scatter_bench 26 100000 > gpurun_out/r2a_scatter_100k.log 2>&1
DNZ_TRACE=1 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2a_bench_trace.json 2> gpurun_out/r2a_bench_trace.err
tail -c 1500 gpurun_out/r2a_bench_trace.json
grep -c superbatch gpurun_out/r2a_bench_trace.err
