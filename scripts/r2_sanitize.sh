# compute-sanitizer is closed on this GPU pool; what stands in for it: the kernels' own invariant checks (-DWAH_TRACE build)
# over a pass that reaches every path, and over the randomised differential test
python scripts/sanitize_small.py 2>&1 | tee gpurun_out/invariants_small.log
WAH_B200_LIB=$PWD/gpu-wah_b200/build_trace/lib/libwah_b200.so python scripts/sanitize_small.py 2>&1 | tee -a gpurun_out/invariants_small.log
WAH_B200_LIB=$PWD/gpu-wah_b200/build_trace/lib/libwah_b200.so timeout 200 python scripts/fuzz_gpu.py 60 4242 2>&1 | tail -2 | tee -a gpurun_out/invariants_small.log
compute-sanitizer --tool memcheck python scripts/sanitize_small.py 2>&1 | tail -2 | tee -a gpurun_out/invariants_small.log
