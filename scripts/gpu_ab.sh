python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-e2e --no-cpu > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
python bench.py --no-e2e --no-cpu --flags 8 > gpurun_out/bench_i_noq.json 2>/dev/null
python bench.py --no-e2e --no-cpu --flags 4 > gpurun_out/bench_i_nohint.json 2>/dev/null
python bench.py --no-e2e --no-cpu --flags 12 > gpurun_out/bench_i_nohint_noq.json 2>/dev/null
python - <<PY
import json
for f in ["bench_i","bench_i_noq","bench_i_nohint","bench_i_nohint_noq"]:
    d=json.load(open("gpurun_out/"+f+".json")); print(f, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
PY
ncu --set full --clock-control none --import-source on -k regex:k_aggregate -s 2 -c 1 -o gpurun_out/prof_agg_r1i -f python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
