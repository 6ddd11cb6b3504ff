"""Runs the reference's own CUDA compress()/decompress() (oracle/_ref/libgpuwah_ref.so, built
untouched for sm_100a) in a process of its own: the reference reads a never-initialised staging
buffer (compress.cu:93, kernels.cu:276) and so must not share an address space -- and with it
recycled device memory -- with anything else.  TEST INFRASTRUCTURE ONLY.

usage: ref_runner.py <libgpuwah_ref.so> compress|decompress|time <in.npy> <out.npy>
"""
import ctypes
import sys

import numpy as np


def load(path):
    lib = ctypes.CDLL(path)
    comp, dec = lib._Z8compressPjyPyPfS1_S1_, lib._Z10decompressPjyPyPfS1_S1_
    pf = ctypes.POINTER(ctypes.c_float)
    for f in (comp, dec):
        f.restype = ctypes.c_void_p
        f.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong), pf, pf, pf]
    return comp, dec


def call(fn, data, times=None):
    n = ctypes.c_ulonglong()
    buf = np.concatenate([data, np.zeros(1, dtype=np.uint32)])   # the reference reads data[n] (kernels.cu:70)
    t = [ctypes.c_float() for _ in range(3)]
    p = fn(buf.ctypes.data, data.size, ctypes.byref(n), ctypes.byref(t[0]), ctypes.byref(t[1]), ctypes.byref(t[2]))
    if not p:
        raise RuntimeError("reference call returned NULL")
    out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), shape=(max(n.value, 1),))[: n.value].copy()
    ctypes.CDLL(None).free(ctypes.c_void_p(p))
    if times is not None:
        times.append([x.value for x in t])
    return out


def main():
    path, what, fin, fout = sys.argv[1:5]
    comp, dec = load(path)
    data = np.load(fin)
    if what == "compress":
        np.save(fout, call(comp, data))
    elif what == "decompress":
        np.save(fout, call(dec, data))
    elif what == "time":
        # the reference benchmark's inner loop (source.cpp:83-126): compress, decompress, compare; 10 reps
        tc, td = [], []
        for _ in range(10):
            c = call(comp, data, tc)
            d = call(dec, c, td)
            assert np.array_equal(d[: data.size], data)
        np.save(fout, np.array([np.median(np.array(tc), axis=0), np.median(np.array(td), axis=0)]))
    else:
        raise SystemExit("unknown command")


if __name__ == "__main__":
    main()
