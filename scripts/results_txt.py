"""results.txt in the reference's column set (source.cpp:38-48,128-138), produced by the drop-in entry points.

The reference's benchmark main sweeps s = 1, 2, 4 ... 256 (dataSize = s * 1024 * 31 * 32 words) x 16 densities
(one set bit in 2^i, i = 1..16) x 10 repetitions and appends one row per (s, i) to results.txt:
original size, compressed size, decompressed size, density index, ratio, and the three timers of compress() and
of decompress() in milliseconds, averaged over the repetitions.  This script writes the same rows with
wah_compress_host / wah_decompress_host, so an old results.txt and a new one can be diffed column by column.
(The reference's main itself links against libwah_b200.so unchanged -- INTEGRATION.md section 1 -- but its rand()
based generator needs minutes per row at the large sizes; here the vectors come from the device generator.)

usage: results_txt.py [--sizes 1,16,256] [--densities 1,4,10,16] [--reps 10] [--out results.txt]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_wah_b200 as wah  # noqa: E402

HEADER = ("Original size [Int] , Compressed size [Int] , Decompressed size [Int] , Density, Compression Ratio, "
          "Compression transfer to device [ms], Compression time [ms], Compression transfer from device [ms], "
          "Decompression transfer to device [ms], Decompression time [ms],Decompression transfer from device [ms]")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,2,4,8,16,32,64,128,256")
    ap.add_argument("--densities", default=",".join(str(i) for i in range(1, 17)))
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default="results.txt")
    a = ap.parse_args()
    rows = [HEADER]
    for s in (int(x) for x in a.sizes.split(",")):
        n = s * 1024 * 31 * 32
        for i in (int(x) for x in a.densities.split(",")):
            data = wah.gen_uniform_device(n, 1.0 / (1 << i), 1337 + i).cpu().numpy().view(np.uint32)
            acc = np.zeros(6)
            for _ in range(a.reps):
                tc, td = {}, {}
                comp = wah.compress(data, wah.WAH_BLOCK1024, tc)
                dec = wah.decompress(comp, td)
                assert dec.size >= n and np.array_equal(dec[:n], data), "data does not match"   # ASSERT, source.cpp:103
                acc += [tc["h2d_ms"], tc["compute_ms"], tc["d2h_ms"], td["h2d_ms"], td["compute_ms"], td["d2h_ms"]]
                csize, dsize = comp.size, dec.size
            acc /= a.reps
            rows.append(f"{n},{csize}, {dsize}, {i}, {csize / n:g}, " + ", ".join(f"{v:g}" for v in acc))
            print(rows[-1], flush=True)
    with open(a.out, "a") as f:
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
