set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not reference_tests_cpp and not reference_kernels and not poisoned" > gpurun_out/pytest_q.log 2>&1; tail -3 gpurun_out/pytest_q.log
grep -q "failed\|error" gpurun_out/pytest_q.log && exit 1
timeout 120 python scripts/fuzz_gpu.py 40 $RANDOM > gpurun_out/dbg_fuzz.log 2>&1; tail -1 gpurun_out/dbg_fuzz.log
grep -q "fuzz ok" gpurun_out/dbg_fuzz.log || exit 1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','compress_gbs','decompress_gbs')})
print('roofline', {k:d['roofline'][k] for k in ('kernel','frac','at','traffic')})
for e in d['sweep']: print(e['density'], e['mode'], round(e['ratio'],5), round(e['compress']['frac'],3), round(e['decompress']['frac'],3))
for k in ('sparse_1gbit','dense_1gbit','bitmap_index'):
    e=d[k]; print(k, round(e['compress']['frac'],3), round(e['decompress']['frac'],3), e.get('value'))
print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
PY
