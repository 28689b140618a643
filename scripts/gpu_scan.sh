python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg2 value', d['value'], 'ms/step', d['ms_per_step'], 'agg ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_scan.csv python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
grep k_tile_scan gpurun_out/launches_scan.csv | awk -F'","' '{print $NF}' | head -4
