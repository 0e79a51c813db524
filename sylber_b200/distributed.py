"""Multi-GPU plumbing for the Segmenter path: one process per GPU, utterances sharded in contiguous blocks, and the
ONE exchange the path has - an all-gather of the fixed-stride segment table.

The reference is single-device (sylber/model/sylber.py:37,54); nothing here exists there.  No stage of the forward
mixes information across utterances, so the data path needs no collective: weights are replicated, every rank runs
`Segmenter` on its shard, and only the small (segment count, segment table) pair is gathered so that each rank
ends up with the boundaries of the whole batch.  Because an utterance's result depends on the batch-wide padded
length T_max (GroupNorm statistics run over padding), ranks first agree on the global T_max.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of utterance indices owned by `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_max_length(local_max: int, device="cpu", group=None) -> int:
    """Batch-wide maximum sample count, so that every rank pads to the T_max a single process would use."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(local_max)
    t = torch.tensor([int(local_max)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def gather_segment_table(seg: torch.Tensor, cnt: torch.Tensor, group=None):
    """All-gather of the per-rank segment table.

    seg (B_local, max_seg, 2) int32 and cnt (B_local,) int32 with the same B_local / max_seg on every rank
    (pad the last shard with zero-count rows).  Returns (world*B_local, max_seg, 2) and (world*B_local,).
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return seg, cnt
    world = dist.get_world_size(group)
    all_seg = torch.empty((world * seg.shape[0],) + tuple(seg.shape[1:]), dtype=seg.dtype, device=seg.device)
    all_cnt = torch.empty((world * cnt.shape[0],), dtype=cnt.dtype, device=cnt.device)
    dist.all_gather_into_tensor(all_cnt, cnt.contiguous(), group=group)
    dist.all_gather_into_tensor(all_seg, seg.contiguous(), group=group)
    return all_seg, all_cnt


def unpack_segment_table(all_seg, all_cnt, n_items, world):
    """Global table -> list of (N_i, 2) int64 arrays in the original utterance order, dropping shard padding."""
    seg = all_seg.cpu().numpy()
    cnt = all_cnt.cpu().numpy()
    per_rank = seg.shape[0] // world
    out = []
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        for k in range(hi - lo):
            row = r * per_rank + k
            out.append(seg[row, :cnt[row]].astype("int64"))
    return out
