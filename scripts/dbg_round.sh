# development helper: host path A/B (spare result buffers)
for nt in 8 16; do
WAH_B200_SPARE_THREADS=$nt timeout 200 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench_spt$nt.json 2>gpurun_out/bench_spt$nt.err; python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['segments_ms'])" gpurun_out/bench_spt$nt.json
done
