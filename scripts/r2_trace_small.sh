# phase timelines at 1 Gbit (fixed costs): decoder and compressor
export WAH_B200_LIB=$PWD/gpu-wah_b200/build_trace/lib/libwah_b200.so
for c in "uniform 0.001 25" "uniform 0.5 25" "clustered 0.5 29"; do set -- $c; python scripts/trace_phases.py --gen $1 --density $2 --log2n $3; done 2>&1 | tee gpurun_out/r2_trace_small.log
python scripts/trace_compress.py 0.001 0 uniform 25 2>&1 | grep -A12 "per-CTA span" | tee -a gpurun_out/r2_trace_small.log
