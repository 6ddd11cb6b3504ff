"""Context figure (not a target): the reference's own CUDA kernels, built untouched for sm_100a with the shuffle
shim (oracle/_ref/libgpuwah_ref.so), on the bench workload at the nearest size the reference is defined for
(n % 992 == 0).  Prints the reference's own three timers (ms) for compress and decompress, medians of 10 runs as in
its benchmark loop (source.cpp:70,83-126), and uncompressed GB/s from its "compute" timer.  TEST INFRASTRUCTURE."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402

lib = os.path.join(ROOT, "oracle", "_ref", "libgpuwah_ref.so")
n = 33_554_400   # 33 825 blocks of 992 words ~ 1 Gbit
density = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
data = datagen.uniform(n, density, 1337)
with tempfile.TemporaryDirectory() as d:
    fin, fout = os.path.join(d, "in.npy"), os.path.join(d, "out.npy")
    np.save(fin, data)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "ref_runner.py"), lib, "time", fin, fout])
    t = np.load(fout)
nb = 4.0 * n
print(f"reference kernels on this GPU, n = {n} words, d = {density}:")
print(f"  compress   H2D {t[0][0]:.3f} ms  compute {t[0][1]:.3f} ms  D2H {t[0][2]:.3f} ms  -> {nb / t[0][1] / 1e6:.1f} GB/s uncompressed (compute timer)")
print(f"  decompress H2D {t[1][0]:.3f} ms  compute {t[1][1]:.3f} ms  D2H {t[1][2]:.3f} ms  -> {nb / t[1][1] / 1e6:.1f} GB/s uncompressed (compute timer)")
