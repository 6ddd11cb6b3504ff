"""Multi-GPU partitioning of the WAH path (one process per GPU, torch.distributed plumbing).

Two workloads shard:

* bitmap index -- independent columns, contiguous blocks of columns per rank, NO collective on
  the data path (an optional all-gather of the per-column lengths gives every rank the table);
* one huge vector -- split at multiples of 992 words (= 1024 groups: a group, word and
  reference-block boundary at once); every rank compresses its range; one all-gather of a
  56-byte record per rank lets every rank compute where its segment lands in the global
  stream and how the fill runs that cross a range boundary are merged (CANONICAL mode; in
  BLOCK1024 mode plain concatenation is already bit-exact).  ``gather_stream`` then
  all-gathers the segments themselves (all-gather-v: grouped sends / receives over NCCL/NVLink).

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
    """Local compress / record on the rank's GPU through the C ABI.  The output buffer, the length word and the
    workspace are kept between calls: the segment ``compress`` returns is a view of that buffer, valid until the next
    ``compress`` of this backend."""

    def __init__(self, device=None):
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._n = -1
        self._out = self._cnt = self._ws = None

    def compress(self, local: torch.Tensor, mode: int):
        n = local.numel()
        if n != self._n:
            cap = wah.max_compressed_words(n)
            self._out = torch.empty(max(cap, 1), dtype=torch.int32, device=self.device)
            self._cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._ws = wah.Workspace.for_compress(n, self.device)
            self._n = n
        wah.compress_device(local, n, self._out, self._out.numel(), self._cnt, self._ws, mode)
        c = int(self._cnt.item())
        if c < 0 or c > self._out.numel():
            raise wah.WahError(2, "the compress kernel reported a failed launch")
        return self._out[:c]

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
    """ONE all-gather of 7 x int64 per rank, one device -> host copy of the result"""
    world = dist.get_world_size(group)
    mine = torch.tensor(rec.as_list(), dtype=torch.int64, device=device)
    gathered = torch.empty(world * mine.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    rows = gathered.view(world, mine.numel()).cpu().tolist()
    return [wah.ShardRecord.from_list(r) for r in rows]


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


def gather_stream(ss: ShardedStream, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather-v of the compressed segments: every rank ends up with the global stream.

    Every rank sends the part of its segment that survives the seam merge straight to its place ``dst[r]`` in every
    other rank's copy of the stream -- one grouped batch of sends and receives (``ncclGroupStart`` / ``ncclSend`` /
    ``ncclRecv`` over NVLink on GPUs), sum of the segment lengths on the wire, nothing padded -- then the few seam
    words of the plan are patched in by one indexed store."""
    world, rank = ss.world, ss.rank
    dev = ss.segment.device
    plan = ss.plan
    lens = [int(r.words) for r in ss.records]
    total = ss.total_words
    if out is None or out.numel() < total:
        out = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    peer = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
    mine = ss.segment[plan["skip"][rank]: lens[rank]]
    ops = []
    for r in range(world):
        body = lens[r] - plan["skip"][r]
        if body <= 0:
            continue
        view = out[plan["dst"][r]: plan["dst"][r] + body]
        if r == rank:
            view.copy_(mine)
        else:
            ops.append(dist.P2POp(dist.irecv, view, peer(r), group))
    if mine.numel() > 0:
        ops += [dist.P2POp(dist.isend, mine, peer(r), group) for r in range(world) if r != rank]
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    # seam words, applied in rank order on the host (a seam may rewrite the previous seam's last word), one store
    patch = {}
    for r in range(world):
        for i, w in enumerate(plan["seam_words"][r]):
            patch[plan["seam_offset"][r] + i] = w - (1 << 32) if w >= (1 << 31) else w
    if patch:
        idx = torch.tensor(list(patch.keys()), dtype=torch.int64, device=dev)
        out[idx] = torch.tensor(list(patch.values()), dtype=torch.int32, device=dev)
    return out[:total]


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
