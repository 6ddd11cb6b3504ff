# One full measurement round on the GPU box: tests, both bench arms, ncu launch list, ncu full capture.
set -x
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.log 2>&1; tail -3 gpurun_out/pytest_gpu_$R.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; tail -1 gpurun_out/smoke_$R.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l_$R.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"wah_(compress|decode)" -s 6 -c 4 -f -o gpurun_out/prof_$R python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f_$R.log 2>&1
cat gpurun_out/bench_$R.json
