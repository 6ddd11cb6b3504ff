"""Multi-GPU partitioning of the WAH path (one process per GPU, torch.distributed plumbing).

Two workloads shard:

* bitmap index -- independent columns, contiguous blocks of columns per rank, NO collective on
  the data path (an optional all-gather of the per-column lengths gives every rank the table);
* one huge vector -- split at multiples of 992 words (= 1024 groups: a group, word and
  reference-block boundary at once); every rank compresses its range; one all-gather of a
  56-byte record per rank lets every rank compute where its segment lands in the global
  stream and how the fill runs that cross a range boundary are merged (CANONICAL mode; in
  BLOCK1024 mode plain concatenation is already bit-exact).  ``gather_stream`` then
  all-gathers the segments themselves over NCCL/NVLink.

The local compressor is injected (``backend``) so the host logic runs unchanged over gloo on
CPU tensors in the tests; the default backend is the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from . import wah

BLOCK_WORDS = 992  # 1024 groups


def column_range(n_cols: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of columns owned by ``rank``."""
    per = (n_cols + world - 1) // world
    lo = min(n_cols, rank * per)
    return lo, min(n_cols, lo + per)


def word_range(n_words: int, rank: int, world: int) -> tuple[int, int]:
    """Range of input words owned by ``rank``: whole 992-word blocks, the last rank takes the tail."""
    blocks = (n_words + BLOCK_WORDS - 1) // BLOCK_WORDS
    per = (blocks + world - 1) // world
    lo = min(n_words, rank * per * BLOCK_WORDS)
    return lo, min(n_words, lo + per * BLOCK_WORDS)


class CudaBackend:
    """Local compress / record on the rank's GPU through the C ABI."""

    def __init__(self, device=None):
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def compress(self, local: torch.Tensor, mode: int):
        n = local.numel()
        cap = wah.max_compressed_words(n)
        out = torch.empty(max(cap, 1), dtype=torch.int32, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        ws = wah.Workspace.for_compress(n, self.device)
        wah.compress_device(local, n, out, cap, cnt, ws, mode)
        c = int(cnt.item())
        return out[:c]

    def record(self, seg: torch.Tensor, groups: int):
        return wah.shard_record_device(seg, seg.numel(), groups)


@dataclass
class ShardedStream:
    """Rank-local view of a range-sharded compressed vector."""

    segment: torch.Tensor            # this rank's compressed words (before seam merging)
    records: list = field(default_factory=list)   # every rank's ShardRecord
    plan: dict = field(default_factory=dict)      # wah.stitch_plan(...) of those records
    mode: int = wah.WAH_BLOCK1024
    rank: int = 0
    world: int = 1

    @property
    def total_words(self) -> int:
        return self.plan["total"]


def _all_gather_records(rec, group, device):
    world = dist.get_world_size(group)
    mine = torch.tensor(rec.as_list(), dtype=torch.int64, device=device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    return [wah.ShardRecord.from_list(t.tolist()) for t in gathered]


def compress_range_sharded(local: torch.Tensor, mode: int = wah.WAH_BLOCK1024, group=None, backend=None) -> ShardedStream:
    """Compress this rank's range of one huge vector and agree on the global layout.

    ``local`` is the rank's slice (``word_range``) of the input.  Exchange: ONE all-gather of a
    7 x int64 record per rank."""
    backend = backend or CudaBackend(local.device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    seg = backend.compress(local, mode)
    rec = backend.record(seg, wah.num_groups(local.numel()))
    records = _all_gather_records(rec, group, local.device)
    plan = wah.stitch_plan(records, mode)
    return ShardedStream(segment=seg, records=records, plan=plan, mode=mode, rank=rank, world=world)


def gather_stream(ss: ShardedStream, group=None) -> torch.Tensor:
    """All-gather the compressed segments and assemble the global stream on every rank.

    Segments are padded to the longest one so a single ``all_gather_into_tensor`` (NCCL over
    NVLink on GPUs) moves them; seams are then patched from the plan."""
    world = ss.world
    dev = ss.segment.device
    lens = [int(r.words) for r in ss.records]
    pad = max(max(lens), 1)
    mine = torch.zeros(pad, dtype=torch.int32, device=dev)
    mine[: ss.segment.numel()] = ss.segment
    allseg = torch.empty(world * pad, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allseg, mine, group=group)
    out = torch.empty(max(ss.total_words, 1), dtype=torch.int32, device=dev)
    plan = ss.plan
    for r in range(world):
        skip, body = plan["skip"][r], lens[r] - plan["skip"][r]
        if body > 0:
            out[plan["dst"][r]: plan["dst"][r] + body] = allseg[r * pad + skip: r * pad + lens[r]]
    for r in range(world):   # in rank order: a seam may overwrite the previous seam's last word
        sw = plan["seam_words"][r]
        if sw:
            vals = torch.tensor([w - (1 << 32) if w >= (1 << 31) else w for w in sw], dtype=torch.int32, device=dev)
            out[plan["seam_offset"][r]: plan["seam_offset"][r] + len(sw)] = vals
    return out[: ss.total_words]


def compress_columns_sharded(local_cols: torch.Tensor, mode: int = wah.WAH_BLOCK1024, group=None,
                             share_lengths: bool = True, backend=None):
    """Bitmap index: ``local_cols`` is this rank's [cols, words_per_col] block (``column_range``).

    Returns (compressed words, local column offsets[cols+1], all ranks' column lengths or None).
    No collective touches the data; ``share_lengths`` all-gathers the per-column lengths."""
    n_cols, wpc = local_cols.shape
    dev = local_cols.device
    if backend is not None:
        out, offs = backend.compress_batch(local_cols, mode)
    else:
        cap = wah.max_compressed_words(wpc) * n_cols
        out = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        offs = torch.zeros(n_cols + 1, dtype=torch.int64, device=dev)
        ws = wah.Workspace.for_compress_batch(n_cols, wpc, dev)
        wah.compress_batch_device(local_cols, n_cols, wpc, local_cols.stride(0), out, cap, offs, ws, mode)
    lengths = None
    if share_lengths and dist.is_initialized():
        world = dist.get_world_size(group)
        mine = (offs[1:] - offs[:-1]).contiguous()
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([n_cols], dtype=torch.int64, device=dev), group=group)
        width = max(int(s.item()) for s in sizes)
        padded = torch.zeros(max(width, 1), dtype=torch.int64, device=dev)
        padded[:n_cols] = mine
        allv = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(allv, padded, group=group)
        lengths = [allv[r][: int(sizes[r].item())] for r in range(world)]
    return out, offs, lengths


def decompress_columns(out: torch.Tensor, offs: torch.Tensor, words_per_col: int, c_total: int | None = None) -> torch.Tensor:
    """Decode this rank's columns again (the result of ``compress_columns_sharded``): [cols, words_per_col].
    Local work only, no collective: ONE launch for all columns.  ``c_total`` = ``offs[-1]`` if the caller knows it
    (saves the device -> host read of that one number)."""
    n_cols = offs.numel() - 1
    dev = out.device
    if n_cols == 0:
        return torch.empty(0, words_per_col, dtype=torch.int32, device=dev)
    if c_total is None:
        c_total = int(offs[-1].item())
    stride = (words_per_col + 1 + 3) // 4 * 4
    back = torch.empty(n_cols * stride, dtype=torch.int32, device=dev)
    info = torch.zeros(3, dtype=torch.int64, device=dev)
    ws = wah.Workspace.for_decompress_batch(n_cols, c_total, words_per_col, dev)
    wah.decompress_batch_device(out, c_total, n_cols, words_per_col, back, stride, words_per_col + 1, info, ws)
    status = int(info[2].item())
    if status:
        raise wah.WahError(5, f"batch decode status {status:#x}")
    return back.view(n_cols, stride)[:, :words_per_col]
