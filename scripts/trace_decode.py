"""Phase timeline of the fused decode kernel (needs a -DWAH_TRACE build: make -C gpu-wah_b200 TRACE=1)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah

n = 1 << 25
dens = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
clustered = len(sys.argv) > 3 and sys.argv[3] == "clustered"
d = wah.gen_clustered_device(n, dens, 1000.0, 1337) if clustered else wah.gen_uniform_device(n, dens, 1337)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_compress(n)
wah.compress_device(d, n, out, cap, cnt, ws, 0)
c = int(cnt.item())
dec = torch.empty(n + 32, dtype=torch.int32, device="cuda")
info = torch.zeros(3, dtype=torch.int64, device="cuda")
wd = wah.Workspace.for_decompress(c, n + 32)
NC = 444
trace = torch.zeros(NC * 64, dtype=torch.int64, device="cuda")
wah.lib.wah_test_set_trace.argtypes = [ctypes.c_void_p]
# a second, different stream: decoding it right before the traced launch warms the instruction cache only
d2 = wah.gen_uniform_device(n, 0.001, 4242)
out2 = torch.empty(cap, dtype=torch.int32, device="cuda")
wah.compress_device(d2, n, out2, cap, cnt, ws, 0)
c2 = int(cnt.item())
dec2 = torch.empty(n + 32, dtype=torch.int32, device="cuda")
wd2 = wah.Workspace.for_decompress(c2, n + 32)
warm = len(sys.argv) > 2 and sys.argv[2] == "warm"
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
for it in range(3):
    flush.zero_()   # 256 MB written: L2 no longer holds the stream
    if warm:
        wah.lib.wah_test_set_trace(None)
        wah.decompress_device(out2, c2, dec2, n + 32, info, wd2)
    trace.zero_()
    wah.lib.wah_test_set_trace(trace.data_ptr())
    wah.decompress_device(out, c, dec, n + 32, info, wd)
    torch.cuda.synchronize()
wah.lib.wah_test_set_trace(None)
t = trace.cpu().numpy().reshape(NC, 64).astype(np.int64)
us = lambda a: a / 1965.0
print("c_words", c, "scan tiles", (c + 2047) // 2048, "output tiles", (n * 32 // 31 + 8191) // 8192)
print("scan phase   us: med %.2f max %.2f" % (np.median(us(t[:, 1] - t[:, 0])), us(t[:, 1] - t[:, 0]).max()))
act = t[:, 4] > 0
print("  pass 1 (start -> published)  us: med %.2f max %.2f" % (np.median(us(t[act, 4] - t[act, 0])), us(t[act, 4] - t[act, 0]).max()))
print("  chained sum                  us: med %.2f max %.2f" % (np.median(us(t[act, 5] - t[act, 4])), us(t[act, 5] - t[act, 4]).max()))
print("  pass 2 (starts)              us: med %.2f max %.2f" % (np.median(us(t[act, 1] - t[act, 5])), us(t[act, 1] - t[act, 5]).max()))
print("global wait  us: med %.2f max %.2f" % (np.median(us(t[:, 2] - t[:, 1])), us(t[:, 2] - t[:, 1]).max()))
print("expand phase us: med %.2f max %.2f" % (np.median(us(t[:, 3] - t[:, 2])), us(t[:, 3] - t[:, 2]).max()))
print("whole CTA    us: med %.2f max %.2f" % (np.median(us(t[:, 3] - t[:, 0])), us(t[:, 3] - t[:, 0]).max()))
tt = t[:, 8:48]
nt = (tt > 0).sum(axis=1)
print("tiles per CTA: min", nt.min(), "max", nt.max())
per = []
for b in range(NC):
    k = nt[b]
    if k >= 3:
        per.append(us(np.diff(tt[b, :k])))
per = np.concatenate(per)
print("per-tile time us: med %.2f p10 %.2f p90 %.2f" % (np.median(per), np.percentile(per, 10), np.percentile(per, 90)))
p1 = us(t[:, 4] - t[:, 0])
order = np.argsort(-p1)
print("slowest pass 1:", [(int(b), round(float(p1[b]), 2)) for b in order[:12]])
print("pass 1 by blockIdx range:", [round(float(np.median(p1[a:a + 37][act[a:a + 37]])), 2) for a in range(0, 407, 37)])
g = t[:, 6]
g = g - g.min()
print("kernel-start globaltimer per CTA (us): p10 %.2f med %.2f p90 %.2f max %.2f" % (np.percentile(g, 10) / 1e3, np.median(g) / 1e3, np.percentile(g, 90) / 1e3, g.max() / 1e3))
print("start time by blockIdx range (us):", [round(float(np.median(g[a:a + 37])) / 1e3, 2) for a in range(0, 444, 37)])
sm = t[:, 7]
print("CTAs per SM: ", np.bincount(np.bincount(sm.astype(int))))
pub = g / 1e3 + us(t[:, 4] - t[:, 0])
print("publish time (start + pass 1) us: med %.2f max %.2f" % (np.median(pub[act]), pub[act].max()))
end = g / 1e3 + us(t[:, 3] - t[:, 0])
sc_end = g / 1e3 + us(t[:, 1] - t[:, 0])
ex_start = g / 1e3 + us(t[:, 2] - t[:, 0])
print("global timeline (us from first CTA start): scan end med %.2f max %.2f | expand start med %.2f max %.2f | CTA end p10 %.2f med %.2f p90 %.2f max %.2f"
      % (np.median(sc_end), sc_end.max(), np.median(ex_start), ex_start.max(), np.percentile(end, 10), np.median(end), np.percentile(end, 90), end.max()))
first_tile = g / 1e3 + us(tt[:, 0] - t[:, 0])
print("first expand tile resolved at (us): med %.2f max %.2f" % (np.median(first_tile), first_tile.max()))
print("end by blockIdx range (us):", [round(float(np.median(end[a:a + 37])), 2) for a in range(0, 444, 37)])
p2w = us(t[:, 50:58] - t[:, 5:6])
if (t[:, 50] > 0).any():
    print("pass 2, offset known -> warp done (us): fastest warp med %.2f, slowest warp med %.2f max %.2f" % (np.median(p2w.min(axis=1)), np.median(p2w.max(axis=1)), p2w.max()))
    print("pass 2, slowest warp done -> scan_body left (us): med %.2f" % np.median(us(t[:, 1:2] - t[:, 50:58]).min(axis=1)))
late = np.argsort(-end)[:6]
for b in late:
    k = nt[b]
    print("late CTA %d: end %.2f tiles %d tile starts (us, global):" % (b, end[b], k), np.round(g[b] / 1e3 + us(tt[b, :k] - t[b, 0]), 1))
    print("     last tile id %d flags %d words %d | loop left at %.2f, bulk stores drained at %.2f" % (t[b, 58], t[b, 59] >> 32, t[b, 59] & 0xFFFFFFFF, g[b] / 1e3 + us(t[b, 60] - t[b, 0]), g[b] / 1e3 + us(t[b, 61] - t[b, 0])))
print("scan tiles: unit", int((t[:, 48] >> 63).sum()), "of", int((t[:, 49] > 0).sum()), "| nsub", int((t[0, 48] >> 48) & 0x7FFF), "| tile 0: groups", int(t[0, 48] & ((1 << 48) - 1)), "words", int(t[0, 49]))
if t[0, 62]:
    print("INVARIANT VIOLATED (DCHK): code", int(t[0, 62]) >> 48, "value", int(t[0, 62]) & ((1 << 48) - 1))
