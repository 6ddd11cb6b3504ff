"""Scan phase vs expand phase of wah_decode_kernel, per CTA (needs the -DWAH_TRACE library:
    make -C gpu-wah_b200 TRACE=1 OBJ=$PWD/gpu-wah_b200/build_trace OUT=$PWD/gpu-wah_b200/build_trace/lib
    WAH_B200_LIB=gpu-wah_b200/build_trace/lib/libwah_b200.so python scripts/trace_phases.py [--gen clustered] [--density 0.5] [--log2n 27])"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gen", default="clustered")
ap.add_argument("--density", type=float, default=0.5)
ap.add_argument("--log2n", type=int, default=27)
a = ap.parse_args()
n = 1 << a.log2n
d = wah.gen_clustered_device(n, a.density, 1000.0, 1337) if a.gen == "clustered" else wah.gen_uniform_device(n, a.density, 1337)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
wah.compress_device(d, n, out, cap, cnt, wah.Workspace.for_compress(n), 0)
c = int(cnt.item())
dec = torch.empty(n + 32, dtype=torch.int32, device="cuda")
info = torch.zeros(3, dtype=torch.int64, device="cuda")
wd = wah.Workspace.for_decompress(c, n + 32)
NC = 444
trace = torch.zeros(NC * 64, dtype=torch.int64, device="cuda")
wah.lib.wah_test_set_trace.argtypes = [ctypes.c_void_p]
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
for it in range(3):
    flush.zero_()
    trace.zero_()
    wah.lib.wah_test_set_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    wah.decompress_device(out, c, dec, n + 32, info, wd)
    e1.record()
    torch.cuda.synchronize()
wah.lib.wah_test_set_trace(None)
assert torch.equal(dec[:n], d)
t = trace.cpu().numpy().reshape(NC, 64).astype(np.int64)
g = (t[:, 6] - t[:, 6].min()) / 1e3   # CTA start, us (globaltimer)
us = lambda x: x / 1965.0
scan_end = g + us(t[:, 1] - t[:, 0])
end = g + us(t[:, 3] - t[:, 0])
act = t[:, 4] > 0
print(f"{a.gen} d={a.density} n=2^{a.log2n} c={c}: launch {e0.elapsed_time(e1) * 1e3:.1f} us")
print("  CTA start           us: med %.1f max %.1f" % (np.median(g), g.max()))
print("  pass 1 published    us: med %.1f max %.1f" % (np.median(g[act] + us(t[act, 4] - t[act, 0])), (g[act] + us(t[act, 4] - t[act, 0])).max()))
print("  tile offset known   us: med %.1f max %.1f" % (np.median(g[act] + us(t[act, 5] - t[act, 0])), (g[act] + us(t[act, 5] - t[act, 0])).max()))
print("  scan phase left     us: med %.1f max %.1f" % (np.median(scan_end), scan_end.max()))
print("  CTA end             us: p10 %.1f med %.1f p90 %.1f max %.1f" % (np.percentile(end, 10), np.median(end), np.percentile(end, 90), end.max()))
if t[0, 62]:
    print("INVARIANT VIOLATED (DCHK): code", int(t[0, 62]) >> 48, "value", int(t[0, 62]) & ((1 << 48) - 1))
