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
    d_info = torch.zeros(3, dtype=torch.int64, device="cuda")
    wd = wah.Workspace.for_decompress(c, n + 40)
    wah.decompress_device(d_out, c, d_dec, n + 40, d_info, wd)
    words, groups, status = d_info.tolist()
    assert status == 0, f"STATUS {status:#x} {name} n={n}"
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


def check_batch(seed, mode):
    # a bitmap-index batch: columns of one length, every kind of content, one compress and ONE decode launch
    global n_cases
    r = np.random.default_rng(seed)
    wpc = int(2 ** r.uniform(3, 17)) + int(r.integers(0, 40))
    n_cols = int(2 ** r.uniform(0, min(11, 23 - np.log2(wpc))))
    cols = np.zeros((n_cols, wpc), dtype=np.uint32)
    for j in range(n_cols):
        k = int(r.integers(0, 5))
        if k == 1:
            cols[j] = 0xFFFFFFFF
        elif k == 2:
            cols[j] = datagen.uniform(wpc, float(10 ** r.uniform(-4, -0.3)), seed + j)
        elif k == 3:
            cols[j] = datagen.clustered(wpc, float(10 ** r.uniform(-3, -0.3)), float(10 ** r.uniform(1.5, 4.5)), seed + j)
        elif k == 4:
            cols[j] = datagen.group_mix(wpc, r.random() * 0.6, r.random() * 0.4, seed + j, run=int(10 ** r.uniform(0, 3)))
    want, offs = orc.compress_batch(cols, mode)
    cap = wah.max_compressed_words(wpc) * n_cols
    d_out = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
    d_offs = torch.full((n_cols + 1,), -1, dtype=torch.int64, device="cuda")
    wah.compress_batch_device(to_dev(cols.reshape(-1)), n_cols, wpc, wpc, d_out, cap, d_offs, wah.Workspace.for_compress_batch(n_cols, wpc), mode)
    assert np.array_equal(d_offs.cpu().numpy().astype(np.uint64), offs), f"BATCH OFFSETS seed={seed} n_cols={n_cols} wpc={wpc} mode={mode}"
    c_total = int(offs[-1])
    assert np.array_equal(d_out[:c_total].cpu().numpy().view(np.uint32), want), f"BATCH COMPRESS seed={seed} n_cols={n_cols} wpc={wpc}"
    stride = (wpc + 1 + 3) // 4 * 4
    d_back = torch.full((n_cols * stride,), -1, dtype=torch.int32, device="cuda")
    d_info = torch.full((3,), -1, dtype=torch.int64, device="cuda")
    wah.decompress_batch_device(d_out, c_total, n_cols, wpc, d_back, stride, wpc + 1, d_info, wah.Workspace.for_decompress_batch(n_cols, c_total, wpc))
    assert d_info.tolist() == [orc.decoded_words(orc.num_groups(wpc)), orc.num_groups(wpc) * n_cols, 0], f"BATCH INFO {d_info.tolist()} seed={seed} n_cols={n_cols} wpc={wpc}"
    back = d_back.cpu().numpy().view(np.uint32).reshape(n_cols, stride)
    assert np.array_equal(back[:, :wpc], cols), f"BATCH DECODE seed={seed} n_cols={n_cols} wpc={wpc} mode={mode}"
    n_cases += 1


while time.time() < t_end:
    kind = rng.integers(0, 7)
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
    elif kind == 6:
        check_batch(seed, mode)
    else:
        nw = min(n, 1 << 19)
        cw, params = random_stream(nw)
        want = orc.decompress(cw)
        if want.size > (1 << 27):
            continue
        d_in = to_dev(cw)
        d_dec = torch.full((want.size + 8,), -1, dtype=torch.int32, device="cuda")
        d_info = torch.zeros(3, dtype=torch.int64, device="cuda")
        wd = wah.Workspace.for_decompress(cw.size, want.size + 8)
        wah.decompress_device(d_in, cw.size, d_dec, want.size + 8, d_info, wd)
        words, groups, status = d_info.tolist()
        assert status == 0, f"STATUS {status:#x} stream nw={nw} params={params}"
        got = d_dec[:words].cpu().numpy().view(np.uint32)
        assert words == want.size and np.array_equal(got, want), f"DECODE MISMATCH stream nw={nw} params={params} seed-state"
        n_cases += 1
print("fuzz ok:", n_cases, "cases in", budget, "s")
