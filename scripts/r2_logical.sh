set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "logical" > gpurun_out/pytest_logical.log 2>&1; tail -25 gpurun_out/pytest_logical.log
python scripts/time_query_ops.py 2>&1 | tee gpurun_out/time_query_ops.log
WAH_B200_LOGICAL_PLAIN=1 python scripts/time_query_ops.py 2>&1 | tee -a gpurun_out/time_query_ops.log
