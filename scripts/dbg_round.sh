# debugging helper: the reference's big tests, fuzz, benches, dense trace
timeout 100 python scripts/fuzz_gpu.py > gpurun_out/dbg_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_fuzz.log; tail -2 gpurun_out/dbg_fuzz.log
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3.log 2>&1; tail -3 gpurun_out/pytest_s3.log
for w in sparse_1gbit dense_1gbit; do
timeout 100 python bench.py --no-e2e --no-cpu-baseline --workload $w > gpurun_out/bench_s3_$w.json 2>gpurun_out/bench_s3_$w.err; python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1],d['ms_per_step'],d['roofline']['compress'],d['roofline']['decompress'])" gpurun_out/bench_s3_$w.json
done
export WAH_B200_LIB=$PWD/gpu-wah_b200/lib_trace/libwah_b200.so
timeout 120 python scripts/trace_decode.py 0.5 warm > gpurun_out/td_dense7.log 2>&1
grep -E "scan tiles|scan phase|pass|chained|global timeline" gpurun_out/td_dense7.log
