# One full measurement round on the GPU box: tests, both bench arms, ncu launch list, ncu full capture,
# the other workloads of profiles/README.md, a results.txt sample.   usage: gpu_round.sh <tag> [quick]
set -x
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.log 2>&1; tail -3 gpurun_out/pytest_gpu_$R.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; tail -1 gpurun_out/smoke_$R.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l_$R.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_$R.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"wah_(compress|decode)" -s 6 -c 4 -f -o gpurun_out/prof_$R python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f_$R.log 2>&1
cat gpurun_out/bench_$R.json
[ "$2" = quick ] && exit 0
B="python bench.py --no-e2e --no-cpu-baseline --steps 100"
$B --workload dense_1gbit > gpurun_out/wl_dense_$R.json 2>/dev/null
$B --workload uniform_32mbit > gpurun_out/wl_c1_$R.json 2>/dev/null
$B --mode canonical > gpurun_out/wl_canon_$R.json 2>/dev/null
for d in 0.0001 0.01 0.1 0.5; do $B --steps 30 --workload clustered_16gbit --density $d > gpurun_out/wl_clu_${d}_$R.json 2>/dev/null; done
python scripts/results_txt.py --sizes 1,32 --densities 1,10 --reps 3 --out gpurun_out/results_$R.txt > /dev/null 2>&1
for f in gpurun_out/wl_*_$R.json; do python -c "
import json,sys;d=json.load(open(sys.argv[1]));r=d['roofline'];print(sys.argv[1],round(d['value']),round(d['ms_per_step']*1e3,1),round(r['compress']['frac'],3),round(r['decompress']['frac'],3))" $f; done
