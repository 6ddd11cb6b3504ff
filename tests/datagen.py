"""Synthetic bitvectors for the tests (numpy; LSB-first bits in 32-bit words)."""
import numpy as np


def pack_bits(bits: np.ndarray) -> np.ndarray:
    n = (bits.size + 31) // 32
    b = np.zeros(n * 32, dtype=np.uint8)
    b[: bits.size] = bits
    return np.packbits(b, bitorder="little").view(np.uint32)


def uniform(n_words: int, density: float, seed: int = 1337) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty(n_words, dtype=np.uint32)
    step = 1 << 16
    for w0 in range(0, n_words, step):
        k = min(step, n_words - w0)
        out[w0: w0 + k] = pack_bits((rng.random(k * 32) < density).astype(np.uint8))
    return out


def clustered(n_words: int, density: float, mean_run: float = 1000.0, seed: int = 1337) -> np.ndarray:
    """two-state Markov chain: 1-runs ~ Geom(mean_run), 0-runs ~ Geom(mean_run (1-d)/d)"""
    rng = np.random.default_rng(seed)
    n_bits = n_words * 32
    l1 = mean_run
    l0 = max(mean_run * (1 - density) / max(density, 1e-12), 1.0)
    pairs = int(n_bits / (l0 + l1) * 1.5) + 16
    z = rng.geometric(1.0 / l0, pairs)
    o = rng.geometric(1.0 / l1, pairs)
    runs = np.stack([z, o], axis=1).reshape(-1)
    vals = np.tile(np.array([0, 1], dtype=np.uint8), pairs)
    bits = np.repeat(vals, runs)[:n_bits]
    if bits.size < n_bits:
        bits = np.concatenate([bits, np.zeros(n_bits - bits.size, dtype=np.uint8)])
    return pack_bits(bits)


def group_mix(n_words: int, p_zero: float, p_one: float, seed: int = 1337, run: int = 1) -> np.ndarray:
    """stream built from 31-bit groups: zero / all-one / random literal, each drawn `run` groups at a time"""
    rng = np.random.default_rng(seed)
    n_groups = (n_words * 32) // 31 + 1
    kinds = np.repeat(rng.random((n_groups + run - 1) // run), run)[:n_groups]
    vals = rng.integers(1, 0x7FFFFFFF, n_groups, dtype=np.uint64)
    vals[kinds < p_zero] = 0
    vals[(kinds >= p_zero) & (kinds < p_zero + p_one)] = 0x7FFFFFFF
    bits = ((vals[:, None] >> np.arange(31, dtype=np.uint64)) & 1).astype(np.uint8).reshape(-1)
    return pack_bits(bits[: n_words * 32])
