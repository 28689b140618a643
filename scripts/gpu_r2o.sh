mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err
echo rc=$?
grep -E "per rank|parity|checksum|rror|leg" gpurun_out/r2o_bench_n8.err | tail -12
