# A/B of the L2 hints (WAH_B200_L2_HINTS: bit 0 decoder, bit 1 compressor):  scripts/gpu_retry.sh 600 -- 'bash scripts/r2_l2ab.sh'
timeout 120 python scripts/fuzz_gpu.py 20 $RANDOM > gpurun_out/dbg_fuzz.log 2>&1; tail -1 gpurun_out/dbg_fuzz.log
grep -q "fuzz ok" gpurun_out/dbg_fuzz.log || exit 1
fmt='import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r["gen"], r["density"], r["n"], "compress", round(r["compress_ms"], 4), round(r["compress_frac"], 3), "decode", round(r["decode_ms"], 4), round(r["decode_frac"], 3))'
for h in 1 3; do echo "== WAH_B200_L2_HINTS=$h"
  for d in 0.5 0.25 0.1 0.01; do WAH_B200_L2_HINTS=$h python scripts/prof_kernels.py --density $d --mode 0 --log2n 29 --reps 7; done | python -c "$fmt"
  WAH_B200_L2_HINTS=$h python scripts/prof_kernels.py --density 0.01 --mode 1 --log2n 29 --reps 7 | python -c "$fmt"
  WAH_B200_L2_HINTS=$h python scripts/prof_kernels.py --gen uniform --density 0.5 --log2n 25 --reps 9 | python -c "$fmt"
  WAH_B200_L2_HINTS=$h python scripts/prof_kernels.py --gen uniform --density 0.05 --log2n 27 --reps 7 | python -c "$fmt"
  WAH_B200_L2_HINTS=$h python scripts/prof_kernels.py --gen uniform --density 0.001 --log2n 25 --reps 9 | python -c "$fmt"
done
