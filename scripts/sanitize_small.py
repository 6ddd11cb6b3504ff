"""A small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): compress + decompress of a few
vectors that reach every expand path (constant, unit, literal scatter, window padded and unpadded, short last tile), a
column batch, and the logical operators, each checked against the oracle.
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import oracle_lib as orc  # noqa: E402

import gpu_wah_b200 as wah  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


# With the -DWAH_TRACE library (WAH_B200_LIB=gpu-wah_b200/build_trace/lib/libwah_b200.so) the kernels' own invariant checks
# (DCHK in wah_decompress.cu: word ranges, ranks, offsets) are compiled in; the first violation ends up in trace[62].
trace = None
if hasattr(wah.lib, "wah_test_set_trace") and "build_trace" in wah.lib_path:
    import ctypes

    trace = torch.zeros(444 * 64, dtype=torch.int64, device="cuda")
    wah.lib.wah_test_set_trace.argtypes = [ctypes.c_void_p]
    wah.lib.wah_test_set_trace(trace.data_ptr())


cases = {
    "zeros": np.zeros(5 * 992 + 7, dtype=np.uint32),
    "dense": datagen.uniform(6 * 992 + 3, 0.5, 1),
    "sparse": datagen.uniform(9 * 992 + 17, 0.001, 2),
    "clustered": datagen.clustered(12 * 992 + 1, 0.3, 300, 3),
    "mixed_dense": datagen.uniform(7 * 992, 0.05, 4),
    "ones_then_data": np.concatenate([np.full(4 * 992, 0xFFFFFFFF, dtype=np.uint32), datagen.clustered(3 * 992 + 5, 0.5, 1000, 5)]),
}
for name, data in cases.items():
    n = data.size
    for mode in (0, 1):
        want = orc.compress(data, mode)
        cap = wah.max_compressed_words(n)
        out = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        wah.compress_device(dev(data), n, out, cap, cnt, wah.Workspace.for_compress(n), mode)
        c = int(cnt.item())
        assert c == want.size and np.array_equal(out[:c].cpu().numpy().view(np.uint32), want), (name, mode)
        dec = torch.full((n + 8,), -1, dtype=torch.int32, device="cuda")
        info = torch.zeros(3, dtype=torch.int64, device="cuda")
        wah.decompress_device(out, c, dec, n + 8, info, wah.Workspace.for_decompress(c, n + 8))
        words = int(info[0].item())
        assert info.tolist()[2] == 0 and np.array_equal(dec[:words].cpu().numpy().view(np.uint32), orc.decompress(want)), (name, mode)
# a column batch
cols = np.stack([datagen.clustered(2 * 992 + 40, 0.2, 200, 10 + j) for j in range(5)])
nc, wpc = cols.shape
cap = wah.max_compressed_words(wpc) * nc
out = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
offs = torch.zeros(nc + 1, dtype=torch.int64, device="cuda")
wah.compress_batch_device(dev(cols), nc, wpc, wpc, out, cap, offs, wah.Workspace.for_compress_batch(nc, wpc), 0)
ct = int(offs[-1].item())
stride = (wpc + 4) // 4 * 4
back = torch.full((nc * stride,), -1, dtype=torch.int32, device="cuda")
info = torch.zeros(3, dtype=torch.int64, device="cuda")
wah.decompress_batch_device(out, ct, nc, wpc, back, stride, wpc + 1, info, wah.Workspace.for_decompress_batch(nc, ct, wpc))
assert np.array_equal(back.view(nc, stride)[:, :wpc].cpu().numpy().view(np.uint32), cols)
# logical operators
a, b = cases["clustered"], datagen.uniform(cases["clustered"].size, 0.01, 9)
ca, cb = orc.compress(a, 0), orc.compress(b, 1)
n = a.size
cap = wah.max_compressed_words(n)
res = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = wah.Workspace.for_logical(n, ca.size, cb.size)
for op, f in {0: lambda x, y: x & y, 1: lambda x, y: x | y, 2: lambda x, y: x ^ y, 3: lambda x, y: x & ~y}.items():
    for mode in (0, 1):
        wah.logical_device(op, dev(ca), ca.size, dev(cb), cb.size, n, res, cap, cnt, ws, mode)
        c = int(cnt.item())
        want = orc.compress(f(a, b), mode)
        assert c == want.size and np.array_equal(res[:c].cpu().numpy().view(np.uint32), want), (op, mode)
if trace is not None:
    torch.cuda.synchronize()
    wah.lib.wah_test_set_trace(None)
    v = int(trace[62].item())
    assert v == 0, f"kernel invariant violated: code {v >> 48}, value {v & ((1 << 48) - 1)}"
    print("kernel invariant checks (DCHK): none violated")
print("sanitize_small ok")
