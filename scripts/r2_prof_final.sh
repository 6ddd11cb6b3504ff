# round-2 evidence: ncu launch list of a short bench run, full captures of both kernels at the sweep's worst and headline points
set -x
python bench.py --steps 3 --warmup 3 --no-sweep --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-sweep --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
for d in 0.5 0.01; do
  python scripts/prof_kernels.py --density $d --log2n 29 --reps 3 > gpurun_out/r2_plain_$d.json 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:wah_decode -c 1 -f -o gpurun_out/r2final_dec_clu_$d python scripts/prof_kernels.py --density $d --log2n 29 --reps 1 --which decode > gpurun_out/ncu_dec_$d.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:wah_compress -s 1 -c 1 -f -o gpurun_out/r2final_cmp_clu_$d python scripts/prof_kernels.py --density $d --log2n 29 --reps 1 --which compress > gpurun_out/ncu_cmp_$d.log 2>&1
done
cat gpurun_out/r2_plain_*.json
# CANONICAL mode at the sweep's worst point
ncu --set full --clock-control none --import-source on -k regex:wah_decode -c 1 -f -o gpurun_out/r2final_decode_canon_0.5 python scripts/prof_kernels.py --density 0.5 --mode 1 --log2n 29 --reps 1 --which decode > gpurun_out/ncu_dec_canon.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wah_compress -s 1 -c 1 -f -o gpurun_out/r2final_compress_canon_0.5 python scripts/prof_kernels.py --density 0.5 --mode 1 --log2n 29 --reps 1 --which compress > gpurun_out/ncu_cmp_canon.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wah_compress -s 1 -c 1 -f -o gpurun_out/r2final_cmp_canon_0.01 python scripts/prof_kernels.py --density 0.01 --mode 1 --log2n 29 --reps 1 --which compress > gpurun_out/ncu_cmp_canon01.log 2>&1
