export WAH_B200_LIB=$PWD/gpu-wah_b200/build_trace/lib/libwah_b200.so
for a in "clustered 0.01 29"; do set -- $a; python scripts/trace_phases.py --gen $1 --density $2 --log2n $3; done 2>&1 | tee gpurun_out/r2_trace.log
