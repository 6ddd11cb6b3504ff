set -x
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -3 gpurun_out/bench_r1d.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r1d.json 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_r1d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l_r1d.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_r1d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wah_ -s 60 -c 3 -f -o gpurun_out/prof_r1d python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f_r1d.log 2>&1
cat gpurun_out/bench_r1d.json
