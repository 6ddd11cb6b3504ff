# development helper: quick parity + bench, then the full round if the quick part is green
set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not reference_tests_cpp and not reference_kernels" > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
grep -q failed gpurun_out/pytest_q.log && exit 1
grep -q passed gpurun_out/pytest_q.log || exit 1
timeout 60 python scripts/fuzz_gpu.py 30 7 > gpurun_out/dbg_fuzz.log 2>&1; tail -1 gpurun_out/dbg_fuzz.log
grep -q "fuzz ok" gpurun_out/dbg_fuzz.log || exit 1
bash scripts/gpu_round.sh r1f4 quick
for w in dense_1gbit; do python bench.py --no-e2e --no-cpu-baseline --steps 100 --workload $w > gpurun_out/wl_dense_r1f4.json 2>/dev/null; done
python bench.py --no-e2e --no-cpu-baseline --steps 30 --workload clustered_16gbit --density 0.1 > gpurun_out/wl_clu_0.1_r1f4.json 2>/dev/null
