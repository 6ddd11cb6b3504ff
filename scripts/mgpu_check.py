"""Multi-GPU parity check, one process per GPU:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py
Range-sharded compress of one vector (NCCL all-gather of the shard records and of the segments) and
column-sharded compress of a bitmap index, both modes, compared with the CPU oracle on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datagen  # noqa: E402
import oracle_lib as orc  # noqa: E402

import gpu_wah_b200 as wah  # noqa: E402
from gpu_wah_b200 import mgpu  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = {
        "zeros": np.zeros(992 * 4000 + 100, dtype=np.uint32),
        "sparse": datagen.uniform(992 * 5000 + 17, 0.001, 5),
        "clustered": datagen.clustered(992 * 6000 + 1, 0.1, 50000, 9),
        "ones_zeros": np.concatenate([np.full(992 * 1500, 0xFFFFFFFF, dtype=np.uint32), np.zeros(992 * 2500 + 5, dtype=np.uint32)]),
    }
    for name, data in cases.items():
        for mode in (wah.WAH_BLOCK1024, wah.WAH_CANONICAL):
            lo, hi = mgpu.word_range(data.size, rank, world)
            local_t = torch.from_numpy(data[lo:hi].view(np.int32).copy()).to(dev)
            ss = mgpu.compress_range_sharded(local_t, mode)
            full = mgpu.gather_stream(ss).cpu().numpy().view(np.uint32)
            want = orc.compress(data, mode)
            good = full.size == want.size and np.array_equal(full, want)
            ok = ok and good
            if rank == 0:
                print(f"range-sharded {name:11s} mode {mode}: {'ok' if good else 'MISMATCH'}  ({want.size} words over {world} ranks)", flush=True)
    cols = np.stack([datagen.uniform(65536, 0.001 * (j + 1), 100 + j) for j in range(4 * world)])
    for mode in (wah.WAH_BLOCK1024, wah.WAH_CANONICAL):
        c0, c1 = mgpu.column_range(cols.shape[0], rank, world)
        mine = torch.from_numpy(cols[c0:c1].view(np.int32).copy()).to(dev)
        out, offs, lengths = mgpu.compress_columns_sharded(mine, mode)
        offs_h = offs.cpu().tolist()
        out_h = out.cpu().numpy().view(np.uint32)
        good = all(np.array_equal(out_h[offs_h[j]: offs_h[j + 1]], orc.compress(cols[c0 + j], mode)) for j in range(c1 - c0))
        good = good and torch.cat(lengths).cpu().tolist() == [orc.compress(c, mode).size for c in cols]
        good = good and torch.equal(mgpu.decompress_columns(out, offs, cols.shape[1]), mine)
        ok = ok and good
        if rank == 0:
            print(f"column-sharded mode {mode}: {'ok' if good else 'MISMATCH'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU CHECK", "PASSED" if int(flag.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
