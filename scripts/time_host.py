"""A/B of the host entry points' result-buffer strategies (WAH_B200_RESULT, wah_host.cu) on one vector:
    python scripts/time_host.py [--log2n 29] [--density 0.01] [--steps 4]
runs itself once per strategy (the switch is read once per process) and prints the three host timers of
wah_compress_host / wah_decompress_host and the end-to-end rate of the pair on a pageable input."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=29)
ap.add_argument("--density", type=float, default=0.01)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--one", action="store_true")
a = ap.parse_args()

if not a.one:
    thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "?"
    print("transparent_hugepage:", thp, " cores:", len(os.sched_getaffinity(0)), flush=True)
    for st, sp in (("0", "0"), ("1", "0"), ("2", "0"), ("2", "1")):
        env = dict(os.environ, WAH_B200_RESULT=st, WAH_B200_SPARSE_COPY=sp)
        subprocess.run([sys.executable, __file__, "--one", "--log2n", str(a.log2n), "--density", str(a.density), "--steps", str(a.steps)], env=env, check=False)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpu_wah_b200 as wah  # noqa: E402

n = 1 << a.log2n
dev = torch.device("cuda", 0)
d = wah.gen_clustered_device(n, a.density, 1000.0, 1337, dev)
h = d.cpu().numpy().view(np.uint32)   # pageable
lib = wah.lib
outp, outn = ctypes.c_void_p(), ctypes.c_uint64()
decp, decn = ctypes.c_void_p(), ctypes.c_uint64()
fl = [ctypes.c_float() for _ in range(6)]


def step(check=False):
    rc = lib.wah_compress_host(h.ctypes.data, n, 0, ctypes.byref(outp), ctypes.byref(outn), ctypes.byref(fl[0]), ctypes.byref(fl[1]), ctypes.byref(fl[2]))
    assert rc == 0
    rc = lib.wah_decompress_host(outp.value, outn.value, ctypes.byref(decp), ctypes.byref(decn), ctypes.byref(fl[3]), ctypes.byref(fl[4]), ctypes.byref(fl[5]))
    assert rc == 0
    if check:
        back = np.frombuffer((ctypes.c_uint32 * n).from_address(decp.value), dtype=np.uint32)
        assert np.array_equal(back, h), "host round trip failed"
    lib.wah_free(outp)
    lib.wah_free(decp)


step(check=True)
step()
t0 = time.perf_counter()
seg = [0.0] * 6
for _ in range(a.steps):
    step()
    for i in range(6):
        seg[i] += fl[i].value
dt = (time.perf_counter() - t0) / a.steps
mv_h, mv_d = ctypes.c_uint64(), ctypes.c_uint64()
lib.wah_host_last_transfer_bytes(ctypes.byref(mv_h), ctypes.byref(mv_d))
print(json.dumps({"strategy": os.environ.get("WAH_B200_RESULT"), "sparse_copy": os.environ.get("WAH_B200_SPARSE_COPY"), "last_d2h_mb": round(mv_d.value / 1e6, 1), "ms_per_step": round(dt * 1e3, 2), "gbs": round(8.0 * n / dt / 1e9, 2),
                  "c_h2d": round(seg[0] / a.steps, 2), "c_d2h": round(seg[2] / a.steps, 2), "d_h2d": round(seg[3] / a.steps, 2),
                  "d_compute": round(seg[4] / a.steps, 2), "d_d2h": round(seg[5] / a.steps, 2)}), flush=True)
