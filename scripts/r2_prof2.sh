set -x
for d in 0.5 0.01; do
ncu --set full --clock-control none --import-source on -k regex:wah_decode -c 1 -f -o gpurun_out/r2b_dec_clu_$d python scripts/prof_kernels.py --density $d --log2n 27 --reps 1 --which decode > gpurun_out/ncu_dec_$d.log 2>&1
done
ls -la gpurun_out/r2b*
