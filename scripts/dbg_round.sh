# development helper: quick parity, A/B bench (tickets vs round robin), full tests, fuzz
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not reference_tests_cpp" > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
grep -q passed gpurun_out/pytest_q.log || exit 1
grep -q failed gpurun_out/pytest_q.log && exit 1
for w in sparse_1gbit dense_1gbit; do
for st in 0 1; do
WAH_B200_STATIC_TILES=$st timeout 100 python bench.py --no-e2e --no-cpu-baseline --workload $w > gpurun_out/bench_t_${w}_$st.json 2>gpurun_out/bench_t_${w}_$st.err; python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1],d['ms_per_step'],d['roofline']['compress']['ms'],d['roofline']['decompress']['ms'])" gpurun_out/bench_t_${w}_$st.json
done
done
timeout 100 python scripts/fuzz_gpu.py > gpurun_out/dbg_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_fuzz.log; tail -2 gpurun_out/dbg_fuzz.log
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s4.log 2>&1; tail -3 gpurun_out/pytest_s4.log
