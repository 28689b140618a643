# 2-GPU: NCCL exchange parity test, then the bench at N=2 (key-partitioned `value` + the un-partitioned exchange leg)
python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-3000
