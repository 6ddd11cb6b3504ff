"""ctypes access to the CPU oracle (oracle/libwah_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libwah_oracle.so"))

BLOCK1024, CANONICAL = 0, 1
_u64, _vp = ctypes.c_uint64, ctypes.c_void_p

for name, res, args in [
    ("wah_oracle_num_groups", _u64, [_u64]),
    ("wah_oracle_group", ctypes.c_uint32, [_vp, _u64, _u64]),
    ("wah_oracle_compress", _u64, [_vp, _u64, ctypes.c_int, _vp]),
    ("wah_oracle_decoded_groups", _u64, [_vp, _u64]),
    ("wah_oracle_decoded_words", _u64, [_u64]),
    ("wah_oracle_decompress", _u64, [_vp, _u64, _vp]),
    ("wah_oracle_canonicalize", _u64, [_vp, _u64, _vp]),
    ("wah_oracle_compress_mt", _u64, [_vp, _u64, ctypes.c_int, _vp, ctypes.c_int]),
    ("wah_oracle_decompress_mt", _u64, [_vp, _u64, _vp, ctypes.c_int]),
    ("wah_oracle_max_threads", ctypes.c_int, []),
    ("wah_oracle_compress_batch", _u64, [_vp, _u64, _u64, ctypes.c_int, _vp, _vp]),
    ("wah_oracle_roundtrip_columns_mt", _u64, [_vp, _u64, _u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    ("wah_oracle_gen_clustered", None, [_vp, _u64, ctypes.c_double, ctypes.c_double, _u64]),
    ("wah_oracle_gen_uniform", None, [_vp, _u64, ctypes.c_double, _u64]),
]:
    f = getattr(_lib, name)
    f.restype, f.argtypes = res, args


def _u32(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint32, a.dtype
    return a


def num_groups(n):
    return int(_lib.wah_oracle_num_groups(n))


def decoded_words(g):
    return int(_lib.wah_oracle_decoded_words(g))


def max_threads():
    return int(_lib.wah_oracle_max_threads())


def compress(data, mode=BLOCK1024, threads=1):
    a = _u32(data)
    out = np.empty(max(num_groups(a.size), 1), dtype=np.uint32)
    if threads == 1:
        c = _lib.wah_oracle_compress(a.ctypes.data, a.size, mode, out.ctypes.data)
    else:
        c = _lib.wah_oracle_compress_mt(a.ctypes.data, a.size, mode, out.ctypes.data, threads)
    return out[:c].copy() if c < out.size // 2 else out[:c]


def decoded_groups(cw):
    a = _u32(cw)
    return int(_lib.wah_oracle_decoded_groups(a.ctypes.data, a.size))


def decompress(cw, threads=1):
    a = _u32(cw)
    words = decoded_words(decoded_groups(a))
    out = np.empty(max(words, 1), dtype=np.uint32)
    if threads == 1:
        n = _lib.wah_oracle_decompress(a.ctypes.data, a.size, out.ctypes.data)
    else:
        n = _lib.wah_oracle_decompress_mt(a.ctypes.data, a.size, out.ctypes.data, threads)
    assert n == words
    return out[:words]


def canonicalize(cw):
    a = _u32(cw)
    out = np.empty(max(a.size, 1), dtype=np.uint32)
    c = _lib.wah_oracle_canonicalize(a.ctypes.data, a.size, out.ctypes.data)
    return out[:c].copy()


def compress_batch(cols, mode=BLOCK1024):
    a = _u32(cols)
    n_cols, wpc = a.shape
    out = np.empty(max(num_groups(wpc) * n_cols, 1), dtype=np.uint32)
    offs = np.zeros(n_cols + 1, dtype=np.uint64)
    c = _lib.wah_oracle_compress_batch(a.ctypes.data, n_cols, wpc, mode, out.ctypes.data, offs.ctypes.data)
    return out[:c].copy(), offs


def logical(op, a, b, groups, mode=BLOCK1024):
    """a op b on the runs of two compressed streams (0 AND, 1 OR, 2 XOR, 3 ANDNOT)"""
    x, y = _u32(a), _u32(b)
    out = np.empty(max(groups, 1), dtype=np.uint32)
    _lib.wah_oracle_logical.restype = ctypes.c_uint64
    _lib.wah_oracle_logical.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                        ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    c = _lib.wah_oracle_logical(op, x.ctypes.data, x.size, y.ctypes.data, y.size, groups, mode, out.ctypes.data)
    return out[:c].copy()


def gen_clustered(n_words, density, mean_run_bits=1000.0, seed=1337, out=None):
    """run-clustered bitvector (two-state Markov chain), packed words written directly (no byte-per-bit array)"""
    out = np.empty(n_words, dtype=np.uint32) if out is None else out
    _lib.wah_oracle_gen_clustered(out.ctypes.data, n_words, density, mean_run_bits, seed)
    return out


def gen_uniform(n_words, density, seed=1337, out=None):
    """i.i.d. Bernoulli(density) bits, packed words written directly"""
    out = np.empty(n_words, dtype=np.uint32) if out is None else out
    _lib.wah_oracle_gen_uniform(out.ctypes.data, n_words, density, seed)
    return out


def roundtrip_columns(cols, mode=BLOCK1024, threads=0, verify=False):
    """compress + decompress of every column of a [n_cols, words_per_col] array, the columns dealt to the threads;
    returns (total compressed words, columns that did not round trip -- 0 unless ``verify``)"""
    a = _u32(cols)
    n_cols, wpc = a.shape
    bad = ctypes.c_uint64(0)
    c = _lib.wah_oracle_roundtrip_columns_mt(a.ctypes.data, n_cols, wpc, mode, threads, 1 if verify else 0, ctypes.byref(bad))
    return int(c), int(bad.value)
