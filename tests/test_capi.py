"""The C-ABI library: loads, exports every declared symbol, host-only entry points work (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gpu_wah_b200 as wah
import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIT31, BIT30, M = 0x80000000, 0x40000000, 0x3FFFFFFF


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "wah_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wah_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(wah.lib, n), n


def test_library_exports_the_reference_entry_points():
    # compress.h:12-18 / decompress.h:11-17, C++ linkage: the symbols tests.o / source.o link against
    assert hasattr(wah.lib, "_Z8compressPjyPyPfS1_S1_")
    assert hasattr(wah.lib, "_Z10decompressPjyPyPfS1_S1_")


@pytest.mark.parametrize("n", [0, 1, 30, 31, 32, 992, 2**20, 2**25, 2**32, 2**36 + 5])
def test_size_arithmetic_matches_oracle(n):
    assert wah.num_groups(n) == orc.num_groups(n) == (32 * n + 30) // 31
    assert wah.max_compressed_words(n) == wah.num_groups(n)
    assert wah.decoded_words(n) == orc.decoded_words(n) == (31 * n + 31) // 32


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(wah.WahError) as e:
        wah.compress(np.zeros(31, dtype=np.uint32))
    assert e.value.code in (2, 3)


def test_argument_validation():
    rc = wah.lib.wah_compress_device(None, 10, 7, None, 0, None, None, 0, None)
    assert rc == 1 and b"mode" in wah.lib.wah_last_error_string()
    rc = wah.lib.wah_compress_device(None, 10, 0, None, 0, None, None, 0, None)
    assert rc == 1


def _record(cw):
    """what wah_shard_record_device computes, restated on a host array"""
    r = wah.wah.ShardRecord()
    r.words, r.groups = cw.size, orc.decoded_groups(cw)
    if cw.size:
        f = int(cw[0])
        if f & BIT31:
            r.lead_type = (f >> 30) & 1
            i = 0
            while i < cw.size and (int(cw[i]) & BIT31) and ((int(cw[i]) >> 30) & 1) == r.lead_type:
                r.lead_groups += int(cw[i]) & M
                i += 1
            r.lead_words = i
        l = int(cw[-1])
        if l & BIT31:
            r.trail_groups, r.trail_type = l & M, (l >> 30) & 1
    return r


def _apply_plan(shards, plan):
    out = np.zeros(plan["total"], dtype=np.uint32)
    for r, s in enumerate(shards):
        body = s[plan["skip"][r]:]
        out[plan["dst"][r]: plan["dst"][r] + body.size] = body
    for r in range(len(shards)):
        sw = plan["seam_words"][r]
        out[plan["seam_offset"][r]: plan["seam_offset"][r] + len(sw)] = sw
    return out


def _f(t, n):
    return BIT31 | (t << 30) | n


STITCH_CASES = {
    "simple": [[5, _f(0, 20)], [_f(0, 30), 7], [9]],
    "chain_through_whole_shards": [[3, _f(1, 10)], [_f(1, 1024)], [_f(1, 1024)], [_f(1, 5), 6, _f(0, 2)], [_f(0, 9)]],
    "type_mismatch": [[_f(0, 10)], [_f(1, 10)], [_f(0, 10)]],
    "empty_shards": [[_f(0, 10)], [], [_f(0, 5), 1], []],
    "overflow": [[1, _f(0, M - 5)], [_f(0, 100), 2]],
    "overflow_multiword_lead": [[_f(0, M), _f(0, 17)], [_f(0, M), _f(0, M), _f(0, 3), 4]],
    "exact_full": [[_f(1, M - 10)], [_f(1, 10)], [_f(1, 1)]],
}


@pytest.mark.parametrize("name", list(STITCH_CASES))
def test_stitch_plan_equals_canonicalising_the_concatenation(name):
    shards = [np.array(s, dtype=np.uint32) for s in STITCH_CASES[name]]
    recs = [_record(s) for s in shards]
    plan = wah.stitch_plan(recs, wah.WAH_CANONICAL)
    got = _apply_plan(shards, plan)
    want = orc.canonicalize(np.concatenate(shards))
    assert np.array_equal(got, want), (got.tolist(), want.tolist())
    # BLOCK1024: plain concatenation
    plan = wah.stitch_plan(recs, wah.WAH_BLOCK1024)
    assert np.array_equal(_apply_plan(shards, plan), np.concatenate(shards))


def test_stitch_plan_on_real_shards():
    import datagen

    data = datagen.clustered(992 * 40, 0.02, 5000, 3)
    for world in (2, 3, 8):
        parts = [data[slice(*wah.mgpu.word_range(data.size, r, world))] for r in range(world)]
        assert sum(p.size for p in parts) == data.size
        for mode in (wah.WAH_BLOCK1024, wah.WAH_CANONICAL):
            shards = [orc.compress(p, mode) for p in parts]
            plan = wah.stitch_plan([_record(s) for s in shards], mode)
            assert np.array_equal(_apply_plan(shards, plan), orc.compress(data, mode))


def test_query_operator_argument_validation():
    # rejected before any CUDA call: works without a device
    assert wah.lib.wah_logical_device(9, None, 0, None, 0, 10, 0, None, 0, None, None, 0, None) == 1
    assert b"operator" in wah.lib.wah_last_error_string()
    assert wah.lib.wah_logical_device(0, None, 0, None, 0, 10, 7, None, 0, None, None, 0, None) == 1     # mode
    assert wah.lib.wah_logical_device(0, None, 0, None, 0, 10, 0, None, 0, None, None, 0, None) == 1     # workspace
    assert wah.lib.wah_popcount_device(None, 5, None, None) == 1
    # the workspace holds both decoded operands plus the scratch of one decode and one compress
    n = 1 << 20
    need = wah.lib.wah_logical_workspace_bytes(n, 1000, 70000)
    assert need >= 2 * 4 * n + wah.lib.wah_decompress_workspace_bytes(70000, n + 8) + wah.lib.wah_compress_workspace_bytes(n)
    assert need < 2 * 4 * n + (4 << 20)
