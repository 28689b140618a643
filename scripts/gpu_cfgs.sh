for c in cfg1 cfg3 cfg5; do
  timeout 600 python bench.py --workload $c --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  echo "== $c rc=$?"; tail -2 gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$c.json")); print("$c", "value", d["value"], "ms/step", d["ms_per_step"], "rows_out", d["rows_out_per_step"], "agg ms", d["roofline"]["avg_launch_ms"], "frac", d["roofline"]["frac"], "B/row", d["roofline"]["algorithmic_bytes_per_row"])
except Exception as e: print("$c failed", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
grep -c k_tile_scan gpurun_out/launches_r1l.csv; grep k_tile_scan gpurun_out/launches_r1l.csv | awk -F'","' '{print $NF}' | head -4
