"""Multi-GPU host logic (gpu-wah_b200/mgpu.py) over gloo, world_size 2 and 3, on CPU tensors.

The local compressor is the injected test backend (the CPU oracle); what is under test is the
sharding arithmetic, the record exchange, the seam plan (wah_stitch_plan in the C ABI, pure host
code) and the assembly of the global stream -- the parts that do not need a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class OracleBackend:
    """compress / record / compress_batch on CPU tensors with the oracle (test infrastructure)."""

    def compress(self, local, mode):
        import oracle_lib as orc

        cw = orc.compress(local.numpy().view(np.uint32), mode)
        return torch.from_numpy(cw.view(np.int32).copy())

    def record(self, seg, groups):
        import gpu_wah_b200 as wah

        cw = seg.numpy().view(np.uint32)
        rec = wah.ShardRecord()
        rec.words, rec.groups = cw.size, groups
        if cw.size:
            first, last = int(cw[0]), int(cw[-1])
            if first >> 31:
                t = (first >> 30) & 1
                k = 0
                while k < cw.size and (int(cw[k]) >> 31) and ((int(cw[k]) >> 30) & 1) == t:
                    rec.lead_groups += int(cw[k]) & 0x3FFFFFFF
                    k += 1
                rec.lead_words, rec.lead_type = k, t
            if last >> 31:
                rec.trail_groups, rec.trail_type = last & 0x3FFFFFFF, (last >> 30) & 1
        return rec

    def compress_batch(self, cols, mode):
        import oracle_lib as orc

        out, offs = orc.compress_batch(cols.numpy().view(np.uint32), mode)
        return torch.from_numpy(out.view(np.int32).copy()), torch.from_numpy(np.asarray(offs, dtype=np.int64))


def _inputs(name):
    import datagen

    B = 992
    if name == "zeros":
        return np.zeros(7 * B + 100, dtype=np.uint32)
    if name == "ones_then_sparse":
        return np.concatenate([np.full(5 * B, 0xFFFFFFFF, dtype=np.uint32), datagen.uniform(6 * B + 3, 0.001, 5)])
    if name == "clustered":
        return datagen.clustered(20 * B + 17, 0.2, 30000, 9)
    if name == "dense":
        return datagen.uniform(9 * B, 0.5, 3)
    raise KeyError(name)


NAMES = ["zeros", "ones_then_sparse", "clustered", "dense"]


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_lib as orc
        from gpu_wah_b200 import mgpu

        ok = True
        for name in NAMES:
            ok = ok and _check_range_sharding(name, mode, rank, world, orc, mgpu)

        # bitmap index: columns dealt in contiguous blocks, no data-path collective
        cols = np.stack([_inputs("clustered")[: 3 * 992 + 5], _inputs("zeros")[: 3 * 992 + 5],
                         _inputs("dense")[: 3 * 992 + 5], _inputs("ones_then_sparse")[: 3 * 992 + 5],
                         _inputs("clustered")[992: 4 * 992 + 5]])
        c0, c1 = mgpu.column_range(cols.shape[0], rank, world)
        mine = torch.from_numpy(cols[c0:c1].view(np.int32).copy())
        out, offs, lengths = mgpu.compress_columns_sharded(mine, mode, backend=OracleBackend())
        for j in range(c1 - c0):
            got = out[int(offs[j]): int(offs[j + 1])].numpy().view(np.uint32)
            ok = ok and np.array_equal(got, orc.compress(cols[c0 + j], mode))
        all_lengths = np.concatenate([l.numpy() for l in lengths])
        ok = ok and all_lengths.tolist() == [orc.compress(c, mode).size for c in cols]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _check_range_sharding(name, mode, rank, world, orc, mgpu):
    data = _inputs(name)
    lo, hi = mgpu.word_range(data.size, rank, world)
    local = torch.from_numpy(data[lo:hi].view(np.int32).copy())
    ss = mgpu.compress_range_sharded(local, mode, backend=OracleBackend())
    full = mgpu.gather_stream(ss).numpy().view(np.uint32)
    want = orc.compress(data, mode)
    ok = full.size == want.size and np.array_equal(full, want)
    # the decoded stream is the input again (zero padded to whole groups)
    dec = orc.decompress(full)
    return bool(ok and np.array_equal(dec[: data.size], data))


@pytest.mark.parametrize("world,mode", [(2, 0), (2, 1), (3, 1)])
def test_range_and_column_sharding_over_gloo(mode, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
