python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg2 value', d['value'], 'ms/step', d['ms_per_step'], 'agg ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'])"
timeout 300 python bench.py --workload cfg5 --rows 268435456 --no-e2e --no-cpu --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg5(268M rows) value', d['value'], 'ms/step', d['ms_per_step'], 'agg ms', d['roofline']['avg_launch_ms'])"
