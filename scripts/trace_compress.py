"""Phase timeline of the compress kernel (needs a -DWAH_TRACE build: make -C gpu-wah_b200 TRACE=1)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah

# usage: trace_compress.py [density] [mode 0|1] [uniform|clustered] [log2n]
dens = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
MODE = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << (int(sys.argv[4]) if len(sys.argv) > 4 else 25)
d = wah.gen_clustered_device(n, dens, 1000.0, 1337) if len(sys.argv) > 3 and sys.argv[3] == "clustered" else wah.gen_uniform_device(n, dens, 1337)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_compress(n)
NC = 296
trace = torch.zeros(NC * 64 * 8, dtype=torch.int64, device="cuda")
wah.lib.wah_test_set_trace.argtypes = [ctypes.c_void_p]
for it in range(3):
    trace.zero_()
    wah.lib.wah_test_set_trace(trace.data_ptr())
    wah.compress_device(d, n, out, cap, cnt, ws, MODE)
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(NC, 64, 8).astype(np.int64)
iters = 14
names = ["tma_issue", "full_done", "classify_done", "agg_seen", "pref_sent", "pref_seen", "emit_done", "gtime"]
def rep(label, a):
    a = a / 1965.0  # us at 1965 MHz
    print(f"{label:34s} med {np.median(a):6.2f} p10 {np.percentile(a,10):6.2f} p90 {np.percentile(a,90):6.2f} us")
for i in (3, 6, 9, 12):
    print("--- iteration", i)
    x = t[:, i, :]
    rep("tma issue -> full seen (worker)", x[:, 1] - x[:, 0])
    rep("classify+scan", x[:, 2] - x[:, 1])
    rep("compact/emit (worker)", x[:, 6] - x[:, 2])
    rep("worker iteration", x[:, 6] - t[:, i - 1, 6])
    rep("worker done -> agg seen (ctl)", x[:, 3] - x[:, 6])
    rep("look-back (agg seen -> pref sent)", x[:, 4] - x[:, 3])
    rep("copy-out (pref sent -> done)", x[:, 5] - x[:, 4])
    rep("control iteration", x[:, 5] - t[:, i - 1, 5])
    rep("worker lead over control (tiles)", 1965.0 * np.array([np.searchsorted(t[b, :iters, 6], x[b, 5]) - i for b in range(NC)]))
last = t[:, :iters, 5].max()
first = t[:, 0, 0].min()
print("per-CTA span us: med %.1f" % np.median((t[:, :iters, 5].max(axis=1) - t[:, 0, 0]) / 1965.0))
print("per-iteration medians (us since the CTA's first TMA issue): full_seen  classify_done  worker_done  agg_seen  pref_sent  copied")
for i in range(iters + 1):
    x = t[:, i, :]
    ok = x[:, 6] > 0
    if ok.sum() == 0:
        continue
    z = lambda col: np.median((x[ok, col] - t[ok, 0, 0]) / 1965.0)
    print(f"  iter {i:2d} ({ok.sum():3d} CTAs): {z(1):7.2f} {z(2):7.2f} {z(6):7.2f} {z(3):7.2f} {z(4):7.2f} {z(5):7.2f}")
print("descriptor re-polls per tile: median per iteration", [int(np.median(t[:, i, 7])) for i in range(iters)], "max", [int(t[:, i, 7].max()) for i in range(iters)])
lb = (t[:, :iters, 4] - t[:, :iters, 3]) / 1965.0
pl = t[:, :iters, 7]
for lo_, hi_ in ((0, 0), (1, 2), (3, 10), (11, 10**9)):
    m = (pl >= lo_) & (pl <= hi_)
    if m.sum():
        print(f"  tiles with {lo_}..{hi_} re-polls: {m.sum():5d}  look-back median {np.median(lb[m]):.2f} us")
