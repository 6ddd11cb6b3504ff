import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah
n = 1 << 25
nbuf = 4
ins = [wah.gen_uniform_device(n, 0.001, 1337 + b) for b in range(nbuf)]
cap = wah.max_compressed_words(n)
out = torch.empty(cap, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_compress(n)
cs = []
for b in range(nbuf):
    wah.compress_device(ins[b], n, out, cap, cnt, ws, 0)
    cs.append(int(cnt.item()))
info = torch.zeros(3, dtype=torch.int64, device="cuda")
wd = wah.Workspace.for_decompress(max(cs), n + 32)
dec = torch.empty(n + 32, dtype=torch.int32, device="cuda")
def step(b):
    wah.compress_device(ins[b], n, out, cap, cnt, ws, 0)
    wah.decompress_device(out, cs[b], dec, n + 32, info, wd)
def run(fn, steps=200):
    for i in range(10): fn(i % nbuf)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps): fn(i % nbuf)
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / steps * 1e3
print("stream launches: %.1f us/step" % run(step))
graphs = []
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for b in range(nbuf):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step(b)
        graphs.append(g)
torch.cuda.synchronize()
print("graph replays  : %.1f us/step" % run(lambda b: graphs[b].replay()))
assert torch.equal(dec[:n], ins[(200 - 1) % nbuf])
