"""Host-side mirror of the C ABI in ``include/wah_b200.h`` (ctypes, no torch types in the ABI).

``compress`` / ``decompress`` take and return host ``uint32`` arrays exactly like the
reference's ``compress()`` / ``decompress()`` (compress.h:12-18, decompress.h:11-17):
same word layout, same WAH format, optional (h2d, compute, d2h) millisecond timings.
The ``*_device`` functions work on buffers that already live in HBM (torch tensors or raw
device pointers) and are asynchronous on the given CUDA stream.

There is deliberately no CPU implementation behind any of these: a missing library or a
missing GPU raises.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np

WAH_BLOCK1024 = 0  # bit-exact to the reference encoder (runs never cross 1024 groups)
WAH_CANONICAL = 1  # maximal runs

WAH_MAX_SEAM_WORDS = 8
# d_out_info[2] of the decoders (include/wah_b200.h)
WAH_STATUS_BAD_WORDS_MASK = 0xFFFFFFFF
WAH_STATUS_TIMEOUT = 1 << 32
WAH_STATUS_BATCH_LENGTH = 1 << 33
WAH_OP_AND, WAH_OP_OR, WAH_OP_XOR, WAH_OP_ANDNOT = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
# WAH_B200_LIB: developer override (scripts/ load a -DWAH_TRACE or experimental build of the same library)
lib_path = os.environ.get("WAH_B200_LIB") or os.path.join(_HERE, "lib", "libwah_b200.so")


class WahError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"wah_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(there is no CPU fallback for the WAH kernels)"
    )
lib = ctypes.CDLL(lib_path)

_u64 = ctypes.c_uint64
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_pf = ctypes.POINTER(ctypes.c_float)


class ShardRecord(ctypes.Structure):
    """``wah_shard_record`` (include/wah_b200.h)."""

    _fields_ = [
        ("words", _u64),
        ("groups", _u64),
        ("lead_groups", _u64),
        ("lead_words", _u64),
        ("trail_groups", _u64),
        ("lead_type", ctypes.c_uint32),
        ("trail_type", ctypes.c_uint32),
    ]

    def as_list(self):
        return [self.words, self.groups, self.lead_groups, self.lead_words, self.trail_groups,
                self.lead_type, self.trail_type]

    @classmethod
    def from_list(cls, v):
        return cls(*[int(x) for x in v])


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("wah_last_error_string", ctypes.c_char_p)
_sig("wah_version", ctypes.c_int)
_sig("wah_num_groups", _u64, _u64)
_sig("wah_max_compressed_words", _u64, _u64)
_sig("wah_decoded_words", _u64, _u64)
_sig("wah_compress_workspace_bytes", _sz, _u64)
_sig("wah_compress_device", ctypes.c_int, _vp, _u64, ctypes.c_int, _vp, _u64, _vp, _vp, _sz, _vp)
_sig("wah_compress_batch_workspace_bytes", _sz, _u64, _u64)
_sig("wah_compress_batch_device", ctypes.c_int, _vp, _u64, _u64, _u64, ctypes.c_int, _vp, _u64, _vp, _vp, _sz, _vp)
_sig("wah_decompress_workspace_bytes", _sz, _u64, _u64)
_sig("wah_decompress_device", ctypes.c_int, _vp, _u64, _vp, _u64, _vp, _vp, _sz, _vp)
_sig("wah_decompress_batch_workspace_bytes", _sz, _u64, _u64, _u64)
_sig("wah_decompress_batch_device", ctypes.c_int, _vp, _u64, _u64, _u64, _vp, _u64, _u64, _vp, _vp, _sz, _vp)
_sig("wah_decoded_size_device", ctypes.c_int, _vp, _u64, _vp, _vp, _sz, _vp)
_sig("wah_compress_host", ctypes.c_int, _vp, _u64, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_u64), _pf, _pf, _pf)
_sig("wah_decompress_host", ctypes.c_int, _vp, _u64, ctypes.POINTER(_vp), ctypes.POINTER(_u64), _pf, _pf, _pf)
_sig("wah_free", None, _vp)
_sig("wah_compress_host_into", ctypes.c_int, _vp, _u64, ctypes.c_int, _vp, _u64, ctypes.POINTER(_u64))
_sig("wah_decompress_host_into", ctypes.c_int, _vp, _u64, _vp, _u64, ctypes.POINTER(_u64))
_sig("wah_host_release", None)
_sig("wah_host_last_transfer_bytes", None, ctypes.POINTER(_u64), ctypes.POINTER(_u64))
_sig("wah_shard_record_device", ctypes.c_int, _vp, _u64, _u64, ctypes.POINTER(ShardRecord), _vp)
_sig("wah_stitch_plan", ctypes.c_int, ctypes.POINTER(ShardRecord), ctypes.c_int, ctypes.c_int,
     ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(_u64), ctypes.POINTER(ctypes.c_uint32),
     ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(_u64))
_sig("wah_test_set_max_launch_tiles", None, _u64)
_sig("wah_test_poison_counter_slots", ctypes.c_int)
_sig("wah_popcount_device", ctypes.c_int, _vp, _u64, _vp, _vp)
_sig("wah_logical_workspace_bytes", _sz, _u64, _u64, _u64)
_sig("wah_logical_device", ctypes.c_int, ctypes.c_int, _vp, _u64, _vp, _u64, _u64, ctypes.c_int, _vp, _u64, _vp, _vp, _sz, _vp)
_sig("wah_container_bytes", _u64, _u64, _u64)
_sig("wah_container_pack", ctypes.c_int, _vp, _u64, ctypes.c_int, _u64, _u64, _vp, _vp)
_sig("wah_container_unpack", ctypes.c_int, _vp, _u64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_u64),
     ctypes.POINTER(_u64), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_u64))
_sig("wah_gen_uniform_device", ctypes.c_int, _vp, _u64, ctypes.c_double, _u64, _vp)
_sig("wah_gen_paint_runs_device", ctypes.c_int, _vp, _u64, _vp, _vp, _u64, _vp)


def _check(rc: int):
    if rc != 0:
        raise WahError(rc, lib.wah_last_error_string().decode())


def num_groups(n_words: int) -> int:
    return int(lib.wah_num_groups(n_words))


def max_compressed_words(n_words: int) -> int:
    return int(lib.wah_max_compressed_words(n_words))


def decoded_words(groups: int) -> int:
    return int(lib.wah_decoded_words(groups))


# ----------------------------------------------------------------------------- host API


def _wrap_malloced(ptr: int, n: int) -> np.ndarray:
    """uint32 view of a malloc()ed result; freed when the array is garbage collected."""
    if n == 0:
        lib.wah_free(ptr)
        return np.empty(0, dtype=np.uint32)
    buf = (ctypes.c_uint32 * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.uint32)
    weakref.finalize(buf, lib.wah_free, ptr)
    return arr


def _host_u32(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint32:
        if a.dtype == np.int32:
            a = a.view(np.uint32)
        else:
            raise TypeError(f"expected a uint32 array, got {a.dtype}")
    return a


def _moved():
    """bytes the last host entry point moved over PCIe, each way (all-zero 4 KiB blocks do not travel)"""
    a, b = _u64(), _u64()
    lib.wah_host_last_transfer_bytes(ctypes.byref(a), ctypes.byref(b))
    return {"h2d_bytes": a.value, "d2h_bytes": b.value}


def compress(data, mode: int = WAH_BLOCK1024, timings: dict | None = None) -> np.ndarray:
    """WAH-compress a host array of 32-bit words (reference: compress(), compress.cu:41-209).

    ``timings``, if given, receives ``h2d_ms`` / ``compute_ms`` / ``d2h_ms`` (the reference's three
    optional float out-params)."""
    a = _host_u32(data)
    out, c = _vp(), _u64()
    t = [ctypes.c_float() for _ in range(3)]
    _check(lib.wah_compress_host(a.ctypes.data, a.size, mode, ctypes.byref(out), ctypes.byref(c),
                                 ctypes.byref(t[0]), ctypes.byref(t[1]), ctypes.byref(t[2])))
    if timings is not None:
        timings.update(h2d_ms=t[0].value, compute_ms=t[1].value, d2h_ms=t[2].value, **_moved())
    return _wrap_malloced(out.value, c.value)


def decompress(data, timings: dict | None = None) -> np.ndarray:
    """Decode a host array of WAH words (reference: decompress(), decompress.cu:18-141).
    Returns ceil(31 G / 32) words, G = number of groups in the stream."""
    a = _host_u32(data)
    out, n = _vp(), _u64()
    t = [ctypes.c_float() for _ in range(3)]
    _check(lib.wah_decompress_host(a.ctypes.data, a.size, ctypes.byref(out), ctypes.byref(n),
                                   ctypes.byref(t[0]), ctypes.byref(t[1]), ctypes.byref(t[2])))
    if timings is not None:
        timings.update(h2d_ms=t[0].value, compute_ms=t[1].value, d2h_ms=t[2].value, **_moved())
    return _wrap_malloced(out.value, n.value)


# --------------------------------------------------------------------------- container (host only)


def container_pack(words, mode: int = WAH_BLOCK1024, words_per_stream: int = 0, stream_offsets=None) -> np.ndarray:
    """Wrap compressed words (one stream, or several back to back with their ``stream_offsets``: columns of a bitmap
    index as ``compress_batch_device`` lays them out, shards of one vector) into a self-describing byte buffer:
    header, offset table, words (``wah_container_pack``)."""
    w = _host_u32(words)
    offs = np.ascontiguousarray([0, w.size] if stream_offsets is None else stream_offsets, dtype=np.uint64)
    n_streams = offs.size - 1
    if n_streams < 1:
        raise WahError(1, "a container holds at least one stream")
    total = int(offs[-1])
    if total != w.size:
        raise WahError(1, "the last stream offset must be the number of words")
    out = np.empty(lib.wah_container_bytes(n_streams, total) // 8 + 1, dtype=np.uint64)   # 8-byte aligned storage
    nbytes = lib.wah_container_bytes(n_streams, total)
    _check(lib.wah_container_pack(out.ctypes.data, nbytes, mode, n_streams, words_per_stream, offs.ctypes.data,
                                  w.ctypes.data))
    return out.view(np.uint8)[:nbytes]


def container_unpack(buf):
    """-> (mode, words_per_stream, stream_offsets, words); validates magic, version, sizes, offsets, checksum."""
    raw = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.view(np.uint8).reshape(-1)
    store = np.empty(raw.size // 8 + 1, dtype=np.uint64)   # an 8-byte aligned copy the returned views refer to
    store.view(np.uint8)[: raw.size] = raw
    mode, ns, wps, tw = ctypes.c_int(), _u64(), _u64(), _u64()
    p_offs, p_words = _vp(), _vp()
    _check(lib.wah_container_unpack(store.ctypes.data, raw.size, ctypes.byref(mode), ctypes.byref(ns), ctypes.byref(wps),
                                    ctypes.byref(p_offs), ctypes.byref(p_words), ctypes.byref(tw)))
    base = store.ctypes.data
    offs = store.view(np.uint8)[p_offs.value - base: p_offs.value - base + 8 * (ns.value + 1)].view(np.uint64)
    words = store.view(np.uint8)[p_words.value - base: p_words.value - base + 4 * tw.value].view(np.uint32)
    return mode.value, wps.value, offs, words


# --------------------------------------------------------------------------- device API


def _ptr(x) -> int:
    if x is None:
        return 0
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x)


def _stream(stream) -> int:
    if stream is None:
        import torch

        return int(torch.cuda.current_stream().cuda_stream)
    if hasattr(stream, "cuda_stream"):
        return int(stream.cuda_stream)
    return int(stream)


class Workspace:
    """Device scratch for the ``*_device`` calls (torch-allocated, reusable across calls)."""

    def __init__(self, nbytes: int, device="cuda"):
        import torch

        self.buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)

    @property
    def nbytes(self) -> int:
        return self.buf.numel()

    def data_ptr(self) -> int:
        return self.buf.data_ptr()

    @classmethod
    def for_compress(cls, n_words: int, device="cuda"):
        return cls(lib.wah_compress_workspace_bytes(n_words), device)

    @classmethod
    def for_compress_batch(cls, n_cols: int, words_per_col: int, device="cuda"):
        return cls(lib.wah_compress_batch_workspace_bytes(n_cols, words_per_col), device)

    @classmethod
    def for_decompress(cls, c_words: int, out_capacity_words: int, device="cuda"):
        return cls(lib.wah_decompress_workspace_bytes(c_words, out_capacity_words), device)

    @classmethod
    def for_logical(cls, n_words: int, ca_words: int, cb_words: int, device="cuda"):
        return cls(lib.wah_logical_workspace_bytes(n_words, ca_words, cb_words), device)

    @classmethod
    def for_decompress_batch(cls, n_cols: int, c_total_words: int, words_per_col: int, device="cuda"):
        return cls(lib.wah_decompress_batch_workspace_bytes(n_cols, c_total_words, words_per_col), device)


def compress_device(d_in, n_words: int, d_out, out_capacity_words: int, d_out_words, workspace,
                    mode: int = WAH_BLOCK1024, stream=None) -> None:
    """Asynchronous single-pass compress of ``n_words`` words resident in HBM.
    ``d_out_words``: device int64/uint64 scalar that receives the compressed length."""
    _check(lib.wah_compress_device(_ptr(d_in), n_words, mode, _ptr(d_out), out_capacity_words,
                                   _ptr(d_out_words), _ptr(workspace), workspace.nbytes, _stream(stream)))


def compress_batch_device(d_in, n_cols: int, words_per_col: int, col_stride_words: int, d_out,
                          out_capacity_words: int, d_col_offsets, workspace, mode: int = WAH_BLOCK1024,
                          stream=None) -> None:
    """Compress ``n_cols`` independent bitmap-index columns; ``d_col_offsets``: device int64[n_cols+1]."""
    _check(lib.wah_compress_batch_device(_ptr(d_in), n_cols, words_per_col, col_stride_words, mode,
                                         _ptr(d_out), out_capacity_words, _ptr(d_col_offsets),
                                         _ptr(workspace), workspace.nbytes, _stream(stream)))


def decompress_device(d_in, c_words: int, d_out, out_capacity_words: int, d_out_info, workspace,
                      stream=None) -> None:
    """Asynchronous decode; ``d_out_info``: device int64[3] receiving (decoded words, decoded groups, status:
    0 or ``WAH_STATUS_*`` bits)."""
    _check(lib.wah_decompress_device(_ptr(d_in), c_words, _ptr(d_out), out_capacity_words,
                                     _ptr(d_out_info), _ptr(workspace), workspace.nbytes, _stream(stream)))


def decompress_batch_device(d_in, c_total_words: int, n_cols: int, words_per_col: int, d_out,
                            out_col_stride_words: int, out_col_words: int, d_out_info, workspace, stream=None) -> None:
    """Decode ``n_cols`` bitmap-index columns laid out back to back as ``compress_batch_device`` writes them, in ONE
    launch; column j goes to ``d_out + j * out_col_stride_words``.  ``d_out_info``: device int64[3] receiving
    (words one column decodes to, groups in the whole stream, status)."""
    _check(lib.wah_decompress_batch_device(_ptr(d_in), c_total_words, n_cols, words_per_col, _ptr(d_out),
                                           out_col_stride_words, out_col_words, _ptr(d_out_info), _ptr(workspace),
                                           workspace.nbytes, _stream(stream)))


def decoded_size_device(d_in, c_words: int, d_out_info, workspace, stream=None) -> None:
    _check(lib.wah_decoded_size_device(_ptr(d_in), c_words, _ptr(d_out_info), _ptr(workspace),
                                       workspace.nbytes, _stream(stream)))


# -------------------------------------------------------------------- range sharding


def popcount_device(d_in, c_words: int, d_bits, stream=None) -> None:
    """``d_bits`` (device int64 scalar) = set bits of the vector the stream stands for, from the stream alone."""
    _check(lib.wah_popcount_device(_ptr(d_in), c_words, _ptr(d_bits), _stream(stream)))


def logical_device(op: int, d_a, ca_words: int, d_b, cb_words: int, n_words: int, d_out, out_capacity_words: int,
                   d_out_words, workspace, mode: int = WAH_BLOCK1024, stream=None) -> None:
    """``out = compress(decompress(a) op decompress(b))`` for two compressed vectors of ``n_words`` words each
    (``WAH_OP_AND`` / ``OR`` / ``XOR`` / ``ANDNOT``); workspace: ``Workspace.for_logical``."""
    _check(lib.wah_logical_device(op, _ptr(d_a), ca_words, _ptr(d_b), cb_words, n_words, mode, _ptr(d_out),
                                  out_capacity_words, _ptr(d_out_words), _ptr(workspace), workspace.nbytes,
                                  _stream(stream)))


def shard_record_device(d_shard, words: int, groups: int, stream=None) -> ShardRecord:
    rec = ShardRecord()
    _check(lib.wah_shard_record_device(_ptr(d_shard), words, groups, ctypes.byref(rec), _stream(stream)))
    return rec


def stitch_plan(records, mode: int):
    """Host-side seam plan for concatenating per-shard streams (``wah_stitch_plan``).

    Returns a dict with per-shard ``skip``, ``dst``, ``seam_offset``, ``seam_words`` (list of
    lists) and the ``total`` length of the global stream."""
    n = len(records)
    recs = (ShardRecord * max(n, 1))(*records)
    skip = (_u64 * max(n, 1))()
    dst = (_u64 * max(n, 1))()
    soff = (_u64 * max(n, 1))()
    scnt = (ctypes.c_uint32 * max(n, 1))()
    swords = (ctypes.c_uint32 * (max(n, 1) * WAH_MAX_SEAM_WORDS))()
    total = _u64()
    _check(lib.wah_stitch_plan(recs, n, mode, skip, dst, soff, scnt, swords, ctypes.byref(total)))
    return {
        "skip": [int(skip[i]) for i in range(n)],
        "dst": [int(dst[i]) for i in range(n)],
        "seam_offset": [int(soff[i]) for i in range(n)],
        "seam_words": [[int(swords[i * WAH_MAX_SEAM_WORDS + k]) for k in range(scnt[i])] for i in range(n)],
        "total": int(total.value),
    }


# ------------------------------------------------------------------------ generators


def gen_uniform_device(n_words: int, density: float, seed: int = 1337, device="cuda", out=None, stream=None):
    """i.i.d. Bernoulli(density) bits (SURVEY.md 8d 'uniform' / 'sparse'), counter based."""
    import torch

    if out is None:
        out = torch.empty(n_words, dtype=torch.int32, device=device)
    _check(lib.wah_gen_uniform_device(_ptr(out), n_words, float(density), seed, _stream(stream)))
    return out


def gen_clustered_device(n_words: int, density: float, mean_run_bits: float = 1000.0, seed: int = 1337,
                         device="cuda", stream=None):
    """Two-state Markov bitvector (SURVEY.md 8d 'clustered'): 1-runs geometric with mean
    ``mean_run_bits``, 0-runs geometric with mean ``mean_run_bits * (1-d)/d``."""
    import torch

    n_bits = n_words * 32
    out = torch.zeros(n_words, dtype=torch.int32, device=device)
    if n_words == 0 or density <= 0.0:
        return out
    l1 = float(mean_run_bits)
    l0 = l1 * (1.0 - density) / density
    pairs = int(n_bits / (l0 + l1) * 1.25) + 64
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    zeros = torch.empty(pairs, dtype=torch.float64, device=device).geometric_(1.0 / max(l0, 1.0), generator=g)
    ones = torch.empty(pairs, dtype=torch.float64, device=device).geometric_(1.0 / max(l1, 1.0), generator=g)
    both = torch.stack([zeros, ones], dim=1).reshape(-1).to(torch.int64)
    ends = torch.cumsum(both, 0)
    starts = (ends - both)[1::2].contiguous()
    lens = both[1::2].contiguous()
    keep = starts < n_bits
    starts, lens = starts[keep].contiguous(), lens[keep].contiguous()
    _check(lib.wah_gen_paint_runs_device(_ptr(out), n_words, _ptr(starts), _ptr(lens), starts.numel(),
                                         _stream(stream)))
    return out
