# phase timeline of the compressor:  make -C gpu-wah_b200 TRACE=1 OBJ=$PWD/gpu-wah_b200/build_trace/obj OUT=$PWD/gpu-wah_b200/build_trace/lib; scripts/gpu_retry.sh 300 -- "bash scripts/r2_trace.sh"
export WAH_B200_LIB=$PWD/gpu-wah_b200/build_trace/lib/libwah_b200.so
for c in "0.01 0" "0.01 1" "0.5 0"; do echo "=== density/mode $c"; python scripts/trace_compress.py $c clustered 27 2>&1 | grep -A10 "iteration 9\|per-CTA span\|re-polls"; done | tee gpurun_out/r2_trace_compress.log
