import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah
n = int(sys.argv[1]); d = float(sys.argv[2]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
x = wah.gen_clustered_device(n, d, 1000.0, 1337)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_compress(n)
wah.compress_device(x, n, out, cap, cnt, ws, mode)
c = int(cnt.item())
print("n", n, "c", c, flush=True)
dec = torch.empty(n + 32, dtype=torch.int32, device="cuda")
info = torch.zeros(2, dtype=torch.int64, device="cuda")
wd = wah.Workspace.for_decompress(c, n + 32)
wah.decompress_device(out, c, dec, n + 32, info, wd)
torch.cuda.synchronize()
print("info", info.tolist(), "equal", bool(torch.equal(dec[:n], x)), flush=True)
