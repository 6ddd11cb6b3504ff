# development round: parity first, then fuzz, then timings
set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not reference_tests_cpp and not reference_kernels and not poisoned" > gpurun_out/pytest_q.log 2>&1; tail -15 gpurun_out/pytest_q.log
grep -q "failed\|error" gpurun_out/pytest_q.log && exit 1
timeout 120 python scripts/fuzz_gpu.py 60 $RANDOM > gpurun_out/dbg_fuzz.log 2>&1; tail -3 gpurun_out/dbg_fuzz.log
grep -q "fuzz ok" gpurun_out/dbg_fuzz.log || exit 1
for d in 0.5 0.25 0.1 0.01 0.001 0.0001; do python scripts/prof_kernels.py --density $d --log2n 29 --reps 5 --which decode; done > gpurun_out/r2_new_c3.jsonl 2>gpurun_out/r2_new_c3.err
python scripts/prof_kernels.py --density 0.0001 --mode 1 --log2n 29 --reps 5 --which decode >> gpurun_out/r2_new_c3.jsonl
python scripts/prof_kernels.py --gen uniform --density 0.5 --log2n 25 --reps 7 --which decode >> gpurun_out/r2_new_c3.jsonl
python scripts/prof_kernels.py --gen uniform --density 0.001 --log2n 25 --reps 7 --which decode >> gpurun_out/r2_new_c3.jsonl
python scripts/prof_kernels.py --gen uniform --density 0.05 --log2n 27 --reps 5 --which decode >> gpurun_out/r2_new_c3.jsonl
cat gpurun_out/r2_new_c3.jsonl
[ "$1" = prof ] || exit 0
for d in 0.5 0.01; do
ncu --set full --clock-control none --import-source on -k regex:wah_decode -c 1 -f -o gpurun_out/r2k_dec_clu_$d python scripts/prof_kernels.py --density $d --log2n 27 --reps 1 --which decode > gpurun_out/ncu_dec_$d.log 2>&1
done
