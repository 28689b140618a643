python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in cfg3 cfg5; do
  timeout 900 python bench.py --workload $c --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$c.json")); print("$c", "value", d["value"], "ms/step", d["ms_per_step"], "rows_out", d["rows_out_per_step"], "agg ms", d["roofline"]["avg_launch_ms"], "launches", d["roofline"]["launches"])
except Exception as e: print("$c failed", e)
PY
done
