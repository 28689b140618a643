# full round-end style run: tests, both bench arms, ncu launch list (shares, not absolutes), one ncu --set full capture of k_aggregate
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
cat gpurun_out/bench_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_aggregate -s 2 -c 1 -o gpurun_out/prof_agg_r1z -f python bench.py --rows 268435456 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
