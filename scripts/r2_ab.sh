set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not reference_tests_cpp and not reference_kernels and not results_txt and not host_entry" > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
timeout 120 python scripts/fuzz_gpu.py 40 $RANDOM > gpurun_out/dbg_fuzz.log 2>&1; tail -1 gpurun_out/dbg_fuzz.log
for d in 0.5 0.1 0.01 0.0001; do python scripts/prof_kernels.py --density $d --log2n 29 --reps 5 --which decode | cut -c1-60,130-; done
python scripts/prof_kernels.py --gen uniform --density 0.001 --log2n 25 --reps 9 --which decode | cut -c1-60,118-
python scripts/prof_kernels.py --gen uniform --density 0.5 --log2n 25 --reps 9 --which decode | cut -c1-60,118-
