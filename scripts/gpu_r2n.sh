mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py -x -q -m gpu 2>&1 | tail -8
for i in 1 2; do
DNZ_BENCH_XSTEP=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2n_bench_n2_$i.json 2> gpurun_out/r2n_bench_n2_$i.err; grep -E "per rank|parity|checksum|rror" gpurun_out/r2n_bench_n2_$i.err | tail -5
done
