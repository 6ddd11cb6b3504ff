# round 2, first GPU call: where does the decoder's time go on run-clustered data (configs[2])?
set -x
nvidia-smi -L; nproc; free -g | head -2
for d in 0.5 0.25 0.1 0.01 0.001 0.0001; do python scripts/prof_kernels.py --density $d --log2n 29 --reps 5; done > gpurun_out/r2_base_c3.jsonl 2>gpurun_out/r2_base_c3.err
python scripts/prof_kernels.py --gen uniform --density 0.5 --log2n 25 --reps 7 >> gpurun_out/r2_base_c3.jsonl
python scripts/prof_kernels.py --gen uniform --density 0.001 --log2n 25 --reps 7 >> gpurun_out/r2_base_c3.jsonl
cat gpurun_out/r2_base_c3.jsonl
for d in 0.5 0.1; do
ncu --set full --clock-control none --import-source on -k regex:wah_decode -c 2 -f -o gpurun_out/r2_dec_clu_$d python scripts/prof_kernels.py --density $d --log2n 27 --reps 2 --which decode > gpurun_out/ncu_dec_$d.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:wah_compress -s 1 -c 1 -f -o gpurun_out/r2_cmp_clu_0.5 python scripts/prof_kernels.py --density 0.5 --log2n 27 --reps 2 --which compress > gpurun_out/ncu_cmp_0.5.log 2>&1
ls -la gpurun_out/*.ncu-rep
