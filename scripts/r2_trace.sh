export WAH_B200_LIB=$PWD/gpu-wah_b200/build_trace/lib/libwah_b200.so
for m in 0 1; do echo "=== mode $m"; python scripts/trace_compress.py 0.01 $m clustered 27 2>&1 | grep -A10 "iteration 9\|per-CTA span\|re-polls"; done | tee gpurun_out/r2_trace_compress.log
