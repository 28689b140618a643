python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for f in 0 16; do python bench.py --workload cfg1 --no-e2e --no-cpu --flags $f 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg1 flags=$f value', d['value'], 'ms/step', d['ms_per_step'], 'agg ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'])"; done
python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg2 value', d['value'], 'ms/step', d['ms_per_step'], 'agg ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'])"
