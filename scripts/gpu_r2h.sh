mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print('value', d['value']/1e9, d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, d['e2e']['ms_per_step'], 'pageable', d['e2e']['pageable'], 'cpu', d['cpu_baseline'], d['parity'])
PY
tail -3 gpurun_out/r2h_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err; tail -c 600 gpurun_out/r2h_ref.json
