mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2
/usr/bin/time -v timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench_n8.json'))
print('N=8 value', d['value']/1e9, d['ms_per_step'], 'frac', d['roofline']['frac'], 'partitioned', d['partitioned']['value']/1e9, d['partitioned']['ms_per_step'])
print('e2e', d['e2e']['value']/1e9 if d.get('e2e') else None, d['parity'])
print('exchange', {k:v for k,v in d['exchange'].items() if k!='what'})
PY
grep -E "bench |Error|error|Elapsed|Maximum resident" gpurun_out/r2k_bench_n8.err | tail -16
