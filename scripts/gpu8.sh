# N-GPU sanity at reduced size: the driver's scaling run must not crash at 8 ranks
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --rows 268435456 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?; tail -4 gpurun_out/bench_n$N.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json")); print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "exchange", d.get("exchange",{}).get("value"), d.get("exchange",{}).get("ms_per_step"))
PY
