set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2.log 2>&1; tail -5 gpurun_out/pytest_gpu_r2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -2 gpurun_out/smoke_r2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -5 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_n1.json 2> gpurun_out/r2_bench_ref_n1.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','compress_gbs','decompress_gbs')})
print('roofline', {k:d['roofline'][k] for k in ('kernel','frac','at','traffic')})
for e in d['sweep']: print(e['density'], e['mode'], round(e['ratio'],5), round(e['compress']['frac'],3), round(e['decompress']['frac'],3))
for k in ('sparse_1gbit','dense_1gbit','bitmap_index'):
    e=d[k]; print(k, round(e['compress']['frac'],3), round(e['decompress']['frac'],3), e.get('value'))
print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
print(open('gpurun_out/r2_bench_ref_n1.json').read()[:600])
PY
