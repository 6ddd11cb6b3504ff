# development round on one B200: quick parity, randomised differential test, kernel timings at 16 Gbit and 1 Gbit
#   scripts/gpu_retry.sh 900 -- 'bash scripts/r2_dev.sh'
set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not reference_tests_cpp and not reference_kernels and not results_txt and not host_entry" > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
grep -q "failed\|error" gpurun_out/pytest_q.log && exit 1
timeout 120 python scripts/fuzz_gpu.py 40 $RANDOM > gpurun_out/dbg_fuzz.log 2>&1; tail -1 gpurun_out/dbg_fuzz.log
grep -q "fuzz ok" gpurun_out/dbg_fuzz.log || exit 1
for m in 0 1; do for d in 0.5 0.25 0.1 0.01 0.001 0.0001; do python scripts/prof_kernels.py --density $d --mode $m --log2n 29 --reps 5; done; done | cut -c1-75,118-
python scripts/prof_kernels.py --gen uniform --density 0.001 --log2n 25 --reps 9 | cut -c1-75,118-
python scripts/prof_kernels.py --gen uniform --density 0.5 --log2n 25 --reps 9 | cut -c1-75,118-
python scripts/prof_kernels.py --gen uniform --density 0.05 --log2n 27 --reps 5 | cut -c1-75,118-
