# compressor timings only (after a change that does not touch the decoder): fuzz, then the density sweep in both modes
#   scripts/gpu_retry.sh 600 -- 'bash scripts/r2_ctime.sh'
timeout 120 python scripts/fuzz_gpu.py 25 $RANDOM > gpurun_out/dbg_fuzz.log 2>&1; tail -1 gpurun_out/dbg_fuzz.log
grep -q "fuzz ok" gpurun_out/dbg_fuzz.log || exit 1
for m in 0 1; do for d in 0.5 0.25 0.1 0.01 0.001 0.0001; do python scripts/prof_kernels.py --density $d --mode $m --log2n 29 --reps 7; done; done | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['density'], r.get('mode'), round(r['compress_ms'], 4), round(r['compress_frac'], 3), round(r['decode_ms'], 4), round(r['decode_frac'], 3))"
python scripts/prof_kernels.py --gen uniform --density 0.5 --log2n 25 --reps 9 | cut -c1-75,118-
python scripts/prof_kernels.py --gen uniform --density 0.05 --log2n 27 --reps 5 | cut -c1-75,118-
