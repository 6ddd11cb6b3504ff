import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah
n = 1 << 25
d = wah.gen_uniform_device(n, 0.001, 1337)
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_compress(n)
wah.compress_device(d, n, out, cap, cnt, ws, 0)
c = int(cnt.item())
info = torch.zeros(3, dtype=torch.int64, device="cuda")
wd = wah.Workspace.for_decompress(c, n + 32)
dec = torch.empty(n + 32, dtype=torch.int32, device="cuda")
flush = torch.empty(64 << 20, dtype=torch.int32, device="cuda")
def timeit(fn, reps=30):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
print("size query (memset + scan kernel) us: median %.1f best %.1f" % timeit(lambda: wah.decoded_size_device(out, c, info, wd)))
print("decode (memset + fused kernel)    us: median %.1f best %.1f" % timeit(lambda: wah.decompress_device(out, c, dec, n + 32, info, wd)))
print("compress (memset + kernel)        us: median %.1f best %.1f" % timeit(lambda: wah.compress_device(d, n, out, cap, cnt, ws, 0)))
e = torch.empty(1 << 25, dtype=torch.int32, device="cuda")
print("torch fill 128 MiB                us: median %.1f best %.1f" % timeit(lambda: e.zero_()))
print("empty (event pair only)           us: median %.1f best %.1f" % timeit(lambda: None))
