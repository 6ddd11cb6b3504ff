# A/B of two builds of the compressor:  scripts/gpu_retry.sh 600 -- 'bash scripts/r2_cab.sh gpu-wah_b200/build_q16/lib/libwah_b200.so'
fmt='import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r["gen"], r["density"], r["n"], "compress", round(r["compress_ms"], 4), round(r["compress_frac"], 3), "decode", round(r["decode_ms"], 4), round(r["decode_frac"], 3))'
for lib in "" "$@"; do echo "== ${lib:-default build}"
  [ -n "$lib" ] && export WAH_B200_LIB=$PWD/$lib
  timeout 120 python scripts/fuzz_gpu.py 10 $RANDOM | tail -1
  for m in 0 1; do for d in 0.5 0.1 0.01 0.0001; do python scripts/prof_kernels.py --density $d --mode $m --log2n 29 --reps 7; done; done | python -c "$fmt"
  python scripts/prof_kernels.py --gen uniform --density 0.001 --log2n 25 --reps 9 | python -c "$fmt"
done
