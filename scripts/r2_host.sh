set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_entry or results_txt or mangled or reference_tests_cpp or golden" > gpurun_out/pytest_host.log 2>&1; tail -3 gpurun_out/pytest_host.log
for d in 0.001 0.01 0.1 0.5; do WAH_B200_SPARSE_COPY=1 python scripts/time_host.py --one --steps 3 --density $d 2>&1 | tail -1 | cut -c1-220; done
