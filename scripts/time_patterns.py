import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah
Z = lambda n: 0x80000000 | n
O = lambda n: 0xC0000000 | n
def stream(pattern, groups_target):
    per = sum((w & 0x3FFFFFFF) if w >> 31 else 1 for w in pattern)
    reps = groups_target // per
    return np.tile(np.array(pattern, dtype=np.uint32), reps)
G = (1 << 25) * 32 // 31
pats = {
    "Z31 L           (sparse-like, 2 words / 32 groups)": [Z(31), 5],
    "Z15 L Z15 L     (4 words / 32 groups)": [Z(15), 5, Z(15), 7],
    "Z31 L O31 L     (one-fills of 31)": [Z(31), 5, O(31), 7],
    "Z62 L O1 L      (one-fills of 1)": [Z(61), 5, O(1), 7],
    "Z287 L O31 L    (clustered d=0.1-like)": [Z(287), 5, O(31), 7],
    "Z31 L O300 L    (long one-fills)": [Z(31), 5, O(300), 7],
    "Z31 O31         (fills only)": [Z(31), O(31)],
    "Z3 L            (2 words / 4 groups: 4096 words per tile)": [Z(3), 5],
}
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
for name, pat in pats.items():
    cw = stream(pat, G)
    d = torch.from_numpy(cw.view(np.int32)).cuda()
    c = cw.size
    n = (G * 31 + 31) // 32 + 64
    dec = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.zeros(3, dtype=torch.int64, device="cuda")
    wd = wah.Workspace.for_decompress(c, n)
    ts = []
    for _ in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); wah.decompress_device(d, c, dec, n, info, wd); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{name:62s} c/n {c / (n - 64):.3f}  {ts[len(ts) // 2]:7.1f} us")
