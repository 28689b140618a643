mkdir -p gpurun_out
DNZ_TRACE=1 DNZ_BENCH_XSTEP=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu --no-e2e --no-parity > gpurun_out/r2m_bench_n2_tr.json 2> gpurun_out/r2m_bench_n2_tr.err; grep -E "per rank" gpurun_out/r2m_bench_n2_tr.err
grep "rank 0 xstep" gpurun_out/r2m_bench_n2_tr.err | sed -n 28,34p
for x in 2; do
DNZ_BENCH_XSTEP=$x timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2m_bench_n2_x$x.json 2> gpurun_out/r2m_bench_n2_x$x.err; grep -E "per rank|parity|checksum" gpurun_out/r2m_bench_n2_x$x.err | tail -5
done
timeout 900 python -m pytest tests/test_gpu_exchange.py -x -q -m gpu 2>&1 | tail -3
