import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_wah_b200 as wah
import oracle_lib as orc
n = int(sys.argv[1]); d = float(sys.argv[2]); mode = 0
x = wah.gen_clustered_device(n, d, 1000.0, 1337)
xh = x.cpu().numpy().view(np.uint32)
want = orc.compress(xh, mode)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_compress(n)
wah.compress_device(x, n, out, cap, cnt, ws, mode)
c = int(cnt.item())
got = out[:c].cpu().numpy().view(np.uint32)
print("compress equal to oracle:", c == want.size and np.array_equal(got, want), c, want.size, flush=True)
counts = np.where(want >> 31, want & 0x3FFFFFFF, 1)
print("max fill", counts.max(), "fills > 4 tiles:", int((counts > 4 * 8192).sum()), "fills > 1 tile:", int((counts > 8192).sum()), flush=True)
info = torch.zeros(2, dtype=torch.int64, device="cuda")
wd = wah.Workspace.for_decompress(c, n + 32)
wah.decoded_size_device(out, c, info, wd)
torch.cuda.synchronize()
print("size query ok", info.tolist(), flush=True)
dec = torch.empty(n + 32, dtype=torch.int32, device="cuda")
import ctypes
trace = torch.zeros(444 * 64, dtype=torch.int64, device="cuda")
wah.lib.wah_test_set_trace.argtypes = [ctypes.c_void_p]
wah.lib.wah_test_set_trace(trace.data_ptr())
wah.decompress_device(out, c, dec, n + 32, info, wd)
try:
    torch.cuda.synchronize()
finally:
    pass
t = trace.cpu().numpy().reshape(444, 64)
v = int(t[:, 62].max()) if True else 0
print("debug check word: code", v >> 48, "val", hex(v & 0xFFFFFFFFFFFF), flush=True)
print("decode ok", info.tolist(), bool(torch.equal(dec[:n], x)), flush=True)
bad = torch.nonzero(dec[:n] != x).flatten().cpu().numpy()
print("mismatching words:", bad.size)
if bad.size:
    tiles = np.unique(bad // 7936)
    print("first bad words", bad[:8], "tiles", tiles[:20], "n bad tiles", tiles.size)
    t0 = int(tiles[0])
    w0 = t0 * 7936
    seg_bad = bad[(bad >= w0) & (bad < w0 + 7936)] - w0
    print("tile", t0, "bad word offsets in tile: min", seg_bad.min(), "max", seg_bad.max(), "count", seg_bad.size)
    print("got", [hex(v & 0xFFFFFFFF) for v in dec[bad[:4]].cpu().tolist()], "want", [hex(v & 0xFFFFFFFF) for v in x[bad[:4]].cpu().tolist()])
    # compressed words that cover that tile
    cnts = np.where(want >> 31, want & 0x3FFFFFFF, 1).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(cnts)])
    g0 = t0 * 8192
    i0 = int(np.searchsorted(offs, g0, side="right") - 1)
    i1 = int(np.searchsorted(offs, g0 + 8192, side="left"))
    print("tile words", i0, "..", i1, "count", i1 - i0, [hex(int(v)) for v in want[i0:min(i1, i0 + 12)]])
if bad.size:
    # describe every bad word of the first bad tile relative to the one-fills that cover it
    for b in seg_bad[:40]:
        bit = int(b) * 32
        g = t0 * 8192 + bit // 31
        i = int(np.searchsorted(offs, g, side="right") - 1)
        wv = int(want[i]); a0 = int(offs[i]) - t0 * 8192; a1 = int(offs[i + 1]) - t0 * 8192
        print("  word", int(b), "covered by stream word", i, hex(wv), "groups [%d,%d) -> words [%d,%d)" % (a0, a1, (31 * a0 + 31) >> 5, (31 * a1) >> 5),
              "got", hex(int(dec[w0 + int(b)].item()) & 0xFFFFFFFF), "even" if (i - (i0 & ~3)) % 2 == 0 else "odd", "idx in round", i - (int(i0) & ~3))
