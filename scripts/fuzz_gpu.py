"""Randomised differential test on the GPU: compress == oracle, decompress(compress(x)) == x, decompress of an
arbitrary valid stream == oracle, over random sizes / generators / modes until the time budget is spent.
usage: fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import oracle_lib as orc  # noqa: E402

import gpu_wah_b200 as wah  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t_end = time.time() + budget
n_cases = 0


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def check(name, data, mode):
    global n_cases
    n = data.size
    want = orc.compress(data, mode)
    d_in = to_dev(data)
    cap = wah.max_compressed_words(n)
    d_out = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ws = wah.Workspace.for_compress(n)
    wah.compress_device(d_in, n, d_out, cap, d_cnt, ws, mode)
    c = int(d_cnt.item())
    got = d_out[:c].cpu().numpy().view(np.uint32)
    assert c == want.size and np.array_equal(got, want), f"COMPRESS MISMATCH {name} n={n} mode={mode}"
    d_dec = torch.full((n + 40,), -1, dtype=torch.int32, device="cuda")
    d_info = torch.zeros(2, dtype=torch.int64, device="cuda")
    wd = wah.Workspace.for_decompress(c, n + 40)
    wah.decompress_device(d_out, c, d_dec, n + 40, d_info, wd)
    words, groups = d_info.tolist()
    assert groups == orc.num_groups(n) and words == orc.decoded_words(groups), f"SIZE MISMATCH {name} n={n}"
    assert torch.equal(d_dec[:n], d_in), f"ROUND TRIP MISMATCH {name} n={n} mode={mode}"
    assert not bool(d_dec[n:words].any()), f"PADDING NOT ZERO {name} n={n}"
    n_cases += 1


def random_stream(n_words):
    p_fill, max_count, p_one = rng.random(), int(10 ** rng.uniform(0, 5.5)), rng.random()
    lit = rng.integers(1, 0x7FFFFFFF, size=n_words, dtype=np.int64).astype(np.uint32)
    is_fill = rng.random(n_words) < p_fill
    cnt = np.minimum(rng.geometric(1.0 / max(max_count / 4.0, 1.0), size=n_words), max_count).astype(np.uint32)
    one = (rng.random(n_words) < p_one).astype(np.uint32)
    return np.where(is_fill, np.uint32(0x80000000) | (one << 30) | cnt, lit).astype(np.uint32), (p_fill, max_count, p_one)


while time.time() < t_end:
    kind = rng.integers(0, 6)
    n = int(2 ** rng.uniform(4, 24)) + int(rng.integers(0, 40))
    mode = int(rng.integers(0, 2))
    seed = int(rng.integers(0, 1 << 30))
    if kind == 0:
        d = float(10 ** rng.uniform(-5, -0.3))
        check(f"uniform d={d:.2e} seed={seed}", datagen.uniform(n, d, seed), mode)
    elif kind == 1:
        d, L = float(10 ** rng.uniform(-4, -0.3)), float(10 ** rng.uniform(1.5, 5.5))
        check(f"clustered d={d:.2e} L={L:.0f} seed={seed}", datagen.clustered(n, d, L, seed), mode)
    elif kind == 2:
        pz, po = rng.random() * 0.6, rng.random() * 0.4
        run = int(10 ** rng.uniform(0, 3.5))
        check(f"group_mix pz={pz:.2f} po={po:.2f} run={run} seed={seed}", datagen.group_mix(n, pz, po, seed, run=run), mode)
    elif kind == 3:
        # long runs of ones and zeros with ragged edges
        bits = np.repeat(rng.integers(0, 2, size=max(n * 32 // 4000, 2)).astype(np.uint8), 4000)[: n * 32]
        bits = np.concatenate([bits, np.zeros(n * 32 - bits.size, dtype=np.uint8)])
        data = np.packbits(bits, bitorder="little").view(np.uint32)
        check(f"blocks4000 seed={seed}", data, mode)
    else:
        nw = min(n, 1 << 19)
        cw, params = random_stream(nw)
        want = orc.decompress(cw)
        if want.size > (1 << 27):
            continue
        d_in = to_dev(cw)
        d_dec = torch.full((want.size + 8,), -1, dtype=torch.int32, device="cuda")
        d_info = torch.zeros(2, dtype=torch.int64, device="cuda")
        wd = wah.Workspace.for_decompress(cw.size, want.size + 8)
        wah.decompress_device(d_in, cw.size, d_dec, want.size + 8, d_info, wd)
        words, groups = d_info.tolist()
        got = d_dec[:words].cpu().numpy().view(np.uint32)
        assert words == want.size and np.array_equal(got, want), f"DECODE MISMATCH stream nw={nw} params={params} seed-state"
        n_cases += 1
print("fuzz ok:", n_cases, "cases in", budget, "s")
