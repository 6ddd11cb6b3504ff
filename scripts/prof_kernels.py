"""Small driver for ncu / timing of the two kernels on one synthetic vector.

    python scripts/prof_kernels.py [--gen clustered|uniform] [--density d] [--log2n k] [--mode 0|1] [--reps r]

Prints the CUDA-event time per launch of wah_compress_kernel and wah_decode_kernel (L2 flushed before every launch when
the vector is smaller than 4x L2) and their fraction of the measured copy bandwidth on 4 (n + c) bytes."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_wah_b200 as wah  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gen", default="clustered")
ap.add_argument("--density", type=float, default=0.5)
ap.add_argument("--log2n", type=int, default=27)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--which", default="both")
a = ap.parse_args()

n = 1 << a.log2n
dev = torch.device("cuda", 0)
d = wah.gen_clustered_device(n, a.density, 1000.0, 1337, dev) if a.gen == "clustered" else wah.gen_uniform_device(n, a.density, 1337, dev)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device=dev)
cnt = torch.zeros(1, dtype=torch.int64, device=dev)
ws = wah.Workspace.for_compress(n, dev)
wah.compress_device(d, n, out, cap, cnt, ws, a.mode)
c = int(cnt.item())
dec = torch.empty(n + 32, dtype=torch.int32, device=dev)
info = torch.zeros(3, dtype=torch.int64, device=dev)
wd = wah.Workspace.for_decompress(c, n + 32, dev)
flush = torch.empty(64 << 20, dtype=torch.int32, device=dev) if n * 4 < (512 << 20) else None
peak = 6547.5
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn):
    ts = []
    for _ in range(a.reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


res = {"gen": a.gen, "density": a.density, "n": n, "c": c, "ratio": c / n, "mode": a.mode}
alg = 4.0 * (n + c)
if a.which in ("both", "compress"):
    t = timed(lambda: wah.compress_device(d, n, out, cap, cnt, ws, a.mode))
    res["compress_ms"] = t
    res["compress_frac"] = alg / (t * 1e-3) / 1e9 / peak
if a.which in ("both", "decode"):
    t = timed(lambda: wah.decompress_device(out, c, dec, n + 32, info, wd))
    res["decode_ms"] = t
    res["decode_frac"] = alg / (t * 1e-3) / 1e9 / peak
    assert torch.equal(dec[:n], d), "round trip failed"
print(json.dumps(res))
