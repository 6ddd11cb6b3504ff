"""Device time of the query operators on two 1 Gbit sparse vectors (d = 0.001 / 0.002), CUDA events, 20 launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah  # noqa: E402

n = 1 << 25
cap = wah.max_compressed_words(n)
ws_c = wah.Workspace.for_compress(n)
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
streams, sizes = [], []
for dens, seed in ((0.001, 1), (0.002, 2)):
    x = wah.gen_uniform_device(n, dens, seed)
    out = torch.empty(cap, dtype=torch.int32, device="cuda")
    wah.compress_device(x, n, out, cap, cnt, ws_c, 0)
    c = int(cnt.item())
    streams.append(out[: c + 8].clone())
    sizes.append(c)
    del x, out
res = torch.empty(cap, dtype=torch.int32, device="cuda")
bits = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_logical(n, sizes[0], sizes[1])
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t = 0.0
    for _ in range(reps):
        flush.zero_()   # 256 MB written: nothing of the operands is left in L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t += a.elapsed_time(b)
    return t / reps * 1e3


for name, op in (("AND", wah.WAH_OP_AND), ("OR", wah.WAH_OP_OR), ("XOR", wah.WAH_OP_XOR), ("ANDNOT", wah.WAH_OP_ANDNOT)):
    us = timed(lambda: wah.logical_device(op, streams[0], sizes[0], streams[1], sizes[1], n, res, cap, cnt, ws, 0))
    print(f"{name:6s} {us:7.1f} us  result {int(cnt.item())} words  ({2 * 4 * n / us / 1e3:.0f} GB/s of uncompressed operand bytes)")
c = int(cnt.item())
us = timed(lambda: wah.popcount_device(res, c, bits))
print(f"popcount of a {c}-word stream: {us:.1f} us, {int(bits.item())} bits set")
