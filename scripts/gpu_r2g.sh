mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_exchange.py -x -q > gpurun_out/r2g_pytest_x.log 2>&1; tail -8 gpurun_out/r2g_pytest_x.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_n2.json'))
print('N=2 value', d['value']/1e9, d['ms_per_step'], 'frac', d['roofline']['frac'], 'partitioned', d['partitioned']['value']/1e9, d['partitioned']['ms_per_step'], d['parity'])
PY
grep -E "Error|error" gpurun_out/r2g_bench_n2.err | tail -5
