# multi-GPU round: parity script, then the contract's bench line at N GPUs (launched the way the driver launches it)
N=${1:-2}
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check_n$N.log 2>&1; tail -12 gpurun_out/mgpu_check_n$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo rc=$?; tail -15 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','compress_gbs','decompress_gbs','n_gpus','scaling')})
print('roofline', {k:d['roofline'][k] for k in ('kernel','frac')})
print('e2e', d['e2e'])
print('range', json.dumps(d['range_128gbit'])[:1500])
PY
