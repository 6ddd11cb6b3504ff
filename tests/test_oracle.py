"""The CPU oracle against the reference's golden vectors and its own invariants (no GPU)."""
import json
import os
import subprocess

import numpy as np
import pytest

import datagen
import oracle_lib as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KATS = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))
BIT31, BIT30 = 0x80000000, 0x40000000


@pytest.mark.parametrize("kat", KATS, ids=[k["kat"] for k in KATS])
def test_oracle_reproduces_golden(kat):
    data = np.array(kat["input_words"], dtype=np.uint32)
    want = np.array(kat["compressed_words"], dtype=np.uint32)
    got = orc.compress(data, orc.BLOCK1024)
    assert np.array_equal(got, want)
    # decode gives the input back (n is a multiple of 31 in every KAT: sizes match exactly)
    assert np.array_equal(orc.decompress(got), data)


def test_kat1_is_the_vector_written_in_the_reference_test():
    # tests.cpp:146: {8, 3|BIT31, 4, 1|BIT31, 2|BIT3130, 24|BIT31}
    kat = KATS[0]
    assert kat["compressed_words"] == [8, 3 | BIT31, 4, 1 | BIT31, 2 | BIT31 | BIT30, 24 | BIT31]
    assert all(k["ref_assert_passed"] for k in KATS[:6])


def _stale_wandering_golden():
    # generateWanderingExpectedData, tests.cpp:66-77 (golden of an older kernel)
    e = np.zeros(93, dtype=np.uint32)
    e[0], e[1] = 1, BIT31 | 31
    for i in range(30):
        e[2 + 3 * i: 5 + 3 * i] = [BIT31 | (i + 1), 1, BIT31 | (30 - i)]
    e[91], e[92] = BIT31 | 32, 1      # index 91 overwrites the loop's last entry, as in the reference
    return e


def test_stale_wandering_golden_is_decode_equivalent():
    kat = next(k for k in KATS if k["kat"] == "KAT-7")
    assert not kat["ref_assert_passed"]
    stale = _stale_wandering_golden()
    assert stale.size == 93 and len(kat["compressed_words"]) == 63
    data = np.array(kat["input_words"], dtype=np.uint32)
    assert np.array_equal(orc.decompress(stale), data)
    # and canonicalising the stale golden gives exactly the oracle's answer
    assert np.array_equal(orc.canonicalize(stale), np.array(kat["compressed_words"], dtype=np.uint32))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_tests_oracle")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_reference_tests_cpp_against_oracle():
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_tests_oracle")], capture_output=True, text=True)
    res = dict(line.split()[1:3] for line in r.stdout.splitlines() if line.startswith("RESULT"))
    for name in ["warpCompressionTest", "blockCompressionTest", "blockMergeTest", "blockMergeWithOnesStartsTest",
                 "blockMergeAlternatingTest", "blockMergeFinalLiterals", "zerosTest"]:
        assert res[name] == "1", name
    assert res["blockMergeWanderingLiterals"] == "0" and res["multiBlockTest"] == "0"   # stale goldens


CASES = [
    ("uniform_0.5", lambda: datagen.uniform(50_000, 0.5, 1)),
    ("uniform_1/16", lambda: datagen.uniform(50_000, 1 / 16, 2)),
    ("sparse", lambda: datagen.uniform(200_000, 0.001, 3)),
    ("sparse_1e-4", lambda: datagen.uniform(200_000, 0.0001, 4)),
    ("clustered_0.5", lambda: datagen.clustered(100_000, 0.5, 1000, 5)),
    ("clustered_0.01", lambda: datagen.clustered(300_000, 0.01, 1000, 6)),
    ("zeros", lambda: np.zeros(10_007, dtype=np.uint32)),
    ("ones", lambda: np.full(992 * 7, 0xFFFFFFFF, dtype=np.uint32)),
    ("mix", lambda: datagen.group_mix(40_000, 0.4, 0.3, 7, run=3)),
]


@pytest.mark.parametrize("name,gen", CASES, ids=[c[0] for c in CASES])
def test_oracle_round_trip_and_mode_relations(name, gen):
    data = gen()
    b = orc.compress(data, orc.BLOCK1024)
    c = orc.compress(data, orc.CANONICAL)
    assert c.size <= b.size
    # BLOCK1024 never emits a run longer than one block, and splits exactly at block boundaries
    fills = b[(b & BIT31) != 0] & 0x3FFFFFFF
    assert fills.size == 0 or fills.max() <= 1024
    assert np.array_equal(orc.canonicalize(b), c)
    assert np.array_equal(orc.canonicalize(c), c)
    for s in (b, c):
        d = orc.decompress(s)
        assert orc.decoded_groups(s) == orc.num_groups(data.size)
        assert d.size == orc.decoded_words(orc.num_groups(data.size))
        assert np.array_equal(d[: data.size], data) and not d[data.size:].any()
    # words of a group sequence never contain a literal that should have been a fill
    lit = c[(c & BIT31) == 0]
    assert not np.any(lit == 0) and not np.any(lit == 0x7FFFFFFF)


@pytest.mark.parametrize("n", [0, 1, 2, 30, 31, 32, 61, 62, 63, 991, 992, 993, 1984, 3000])
def test_oracle_tail_sizes(n):
    data = datagen.uniform(n, 0.3, 40 + n)
    for mode in (orc.BLOCK1024, orc.CANONICAL):
        s = orc.compress(data, mode)
        assert orc.decoded_groups(s) == orc.num_groups(n) == (32 * n + 30) // 31
        d = orc.decompress(s)
        assert d.size in (n, n + 1)          # zero-padded last group can spill one extra zero word
        assert np.array_equal(d[:n], data) and not d[n:].any()


def test_oracle_group_definition():
    # tests.cpp:94-97: expected[i] = 0x7FFFFFFF & ((data[i] << i) | data[i-1] >> (32-i))
    data = np.arange(1, 32, dtype=np.uint32) * np.uint32(0x01020305)
    for i in range(32):
        lo = int(data[i]) << i if i < 31 else 0
        hi = int(data[i - 1]) >> (32 - i) if i > 0 else 0
        assert orc._lib.wah_oracle_group(data.ctypes.data, 31, i) == (lo | hi) & 0x7FFFFFFF


@pytest.mark.parametrize("threads", [2, 3, 8])
@pytest.mark.parametrize("mode", [orc.BLOCK1024, orc.CANONICAL])
def test_oracle_mt_equals_sequential(threads, mode):
    for data in (datagen.uniform(70_000, 0.001, 9), datagen.clustered(90_001, 0.05, 3000, 10),
                 np.zeros(992 * 50 + 5, dtype=np.uint32), datagen.uniform(30_000, 0.5, 11)):
        s = orc.compress(data, mode)
        assert np.array_equal(orc.compress(data, mode, threads=threads), s)
        assert np.array_equal(orc.decompress(s, threads=threads), orc.decompress(s))


def test_oracle_canonical_splits_runs_at_the_counter_limit():
    M = 0x3FFFFFFF
    cw = np.array([BIT31 | M, BIT31 | M, BIT31 | 5, BIT31 | BIT30 | 7, BIT31 | BIT30 | M, 9], dtype=np.uint32)
    got = orc.canonicalize(cw)
    assert got.tolist() == [BIT31 | M, BIT31 | M, BIT31 | 5, BIT31 | BIT30 | M, BIT31 | BIT30 | 7, 9]
    assert orc.decoded_groups(got) == orc.decoded_groups(cw)


# --------------------------------------------------------------------------- query operators on the runs themselves

_NP_OPS = {0: lambda a, b: a & b, 1: lambda a, b: a | b, 2: lambda a, b: a ^ b, 3: lambda a, b: a & ~b}


@pytest.mark.parametrize("mode", [orc.BLOCK1024, orc.CANONICAL])
@pytest.mark.parametrize("name,n,gen_a,gen_b", [
    ("sparse_x_sparse", 992 * 40 + 5, lambda n: datagen.uniform(n, 0.001, 1), lambda n: datagen.uniform(n, 0.002, 2)),
    ("clustered_x_dense", 3 * 7936 + 100, lambda n: datagen.clustered(n, 0.3, 1000, 3), lambda n: datagen.uniform(n, 0.5, 4)),
    ("ones_x_zeros", 2 * 7936, lambda n: np.full(n, 0xFFFFFFFF, dtype=np.uint32), lambda n: np.zeros(n, dtype=np.uint32)),
    ("ones_x_clustered", 5000, lambda n: np.full(n, 0xFFFFFFFF, dtype=np.uint32), lambda n: datagen.clustered(n, 0.5, 300, 9)),
    ("clustered_x_clustered", 1 << 18, lambda n: datagen.clustered(n, 0.05, 1000, 5), lambda n: datagen.clustered(n, 0.2, 300, 6)),
    ("tiny", 7, lambda n: datagen.uniform(n, 0.5, 7), lambda n: datagen.uniform(n, 0.5, 8)),
    ("one_word", 1, lambda n: np.array([0x80000001], dtype=np.uint32), lambda n: np.array([0xFFFFFFFF], dtype=np.uint32)),
])
def test_logical_on_runs_equals_compress_of_the_combined_vector(name, n, gen_a, gen_b, mode):
    """The compressed-domain operator (specification of the next kernel) against the plain route: decode both,
    combine word by word, encode.  The operands may come from either encoder mode."""
    a, b = gen_a(n), gen_b(n)
    ca, cb = orc.compress(a, mode), orc.compress(b, 1 - mode)
    for op, f in _NP_OPS.items():
        want = orc.compress(f(a, b), mode)
        got = orc.logical(op, ca, cb, orc.num_groups(n), mode)
        assert np.array_equal(got, want), (name, op)


def test_logical_on_runs_zero_extends_a_short_operand():
    a = datagen.clustered(4000, 0.4, 200, 11)
    b = datagen.uniform(1000, 0.3, 12)
    b_ext = np.concatenate([b, np.zeros(3000, dtype=np.uint32)])
    ca, cb = orc.compress(a, 1), orc.compress(b, 1)
    # (b's own last group is zero padded, exactly as the zero extension continues it)
    for op, f in _NP_OPS.items():
        assert np.array_equal(orc.logical(op, ca, cb, orc.num_groups(4000), 1), orc.compress(f(a, b_ext), 1)), op
