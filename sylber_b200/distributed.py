"""Multi-GPU plumbing for the Segmenter path: one process per GPU, utterances sharded in contiguous blocks, and the
ONE exchange the path has - an all-gather of the fixed-stride segment table.

The reference is single-device (sylber/model/sylber.py:37,54); nothing here exists there.  No stage of the forward
mixes information across utterances, so the data path needs no collective: weights are replicated, every rank runs
`Segmenter` on its shard, and only the small (segment count, segment table) pair is gathered so that each rank
ends up with the boundaries of the whole batch.  Because an utterance's result depends on the batch-wide padded
length T_max (GroupNorm statistics run over padding), ranks first agree on the global T_max.
"""
from __future__ import annotations

import numpy as np
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


LAST_TIMING = {}     # host-side phase times of the most recent segment_sharded fast-path call (tools/shard_probe.py)


def _comm_device(segmenter, group):
    """NCCL moves device tensors, every other backend (gloo in the CPU tests) host tensors."""
    backend = dist.get_backend(group)
    return torch.device(segmenter.device) if "nccl" in str(backend) else torch.device("cpu")


def segment_sharded(segmenter, wav_file=None, wav=None, in_second=True, group=None, local_input=False, pad_to=None,
                    gather_features=False, per_rank=None, hidden_to="host"):
    """`Segmenter.__call__` for a LIST of utterances sharded over the ranks of `group` (one process per GPU).

    Every rank calls this collectively.  With `local_input=False` (default) every rank passes the same global list and
    runs the contiguous block `shard_range(len(list), rank, world)`; with `local_input=True` each rank passes only its
    own clips (global order = rank order).  The forward has no cross-utterance dependency, so the data path needs no
    collective; what is exchanged is (1) with local input and no `pad_to`, the batch-wide maximum length - an
    utterance's result depends on the padded length, so every rank pads to the T_max the single call
    would use - and (2) the segment table: ONE all-gather of the fixed-stride (utterances, 1 + T, 2) int32 block the forward
    left on the device (row 0 carries the segment count; ~130 KB per rank at 32 x 10 s, over NCCL / NVSwitch), enqueued
    behind the sub-batch streams while the host still collects the last hidden states, then one device->host copy.
    (With `gather_features=True`, or a segmenter without device tables, counts, table and features travel as three
    host-packed all-gathers instead.)

    `hidden_to`: where the rank's own hidden states go - "host" (NumPy, default), "device" (torch CUDA tensors: they stay
    sharded ON the GPUs, which is what a multi-GPU pipeline that consumes segments / features wants: the host link is
    shared by all ranks of a box and 49 MB per 32 x 10 s per rank is what saturates it) or None.

    `per_rank` (with `local_input=True`): an upper bound on the number of utterances any rank passes, known to the caller
    (e.g. the per-GPU batch size).  It saves the collective that otherwise exchanges the list lengths on every call; the
    actual lengths then travel inside the segment-table block.

    Returns the list of result dicts of the WHOLE batch in global order on every rank: `segments` for every utterance
    (bit-identical to what one process calling `segmenter(wav=whole_list)` returns - tests/test_gpu_sharded.py);
    `segment_features` and `hidden_states` for the rank's own utterances and None for the others, unless
    `gather_features=True`, which also all-gathers the (N, 768) segment features.  Hidden states stay sharded."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = segmenter(wav_file=wav_file, wav=wav, in_second=in_second, pad_to=pad_to,
                        **({"hidden_to": hidden_to} if hidden_to != "host" else {}))
        return out if isinstance(out, list) else [out]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    items = wav_file if wav_file is not None else wav
    if not isinstance(items, (list, tuple)):
        raise TypeError("segment_sharded shards a list of utterances; pass a list")
    items = list(items)
    dev = _comm_device(segmenter, group)
    sizes_in_block = False
    if local_input and per_rank is not None and hasattr(segmenter, "call_with_tables") and not gather_features:
        if len(items) > per_rank:
            raise ValueError(f"per_rank={per_rank} but this rank passes {len(items)} utterances")
        sizes = None                      # read from the gathered block (header of every rank's first row)
        sizes_in_block = True
        mine = items
    elif local_input:
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        mine_n = torch.tensor([len(items)], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, mine_n, group=group)
        sizes = [int(x) for x in sizes.cpu().tolist()]
        mine = items
    else:
        sizes = [shard_range(len(items), r, world)[1] - shard_range(len(items), r, world)[0] for r in range(world)]
        lo, hi = shard_range(len(items), rank, world)
        mine = items[lo:hi]
    # the padded length of the whole call
    if wav_file is not None:
        rows, _ = segmenter._prepare(mine, None) if mine else ([], True)
        local_kw = {"wav": rows}
        known_all = False
    else:
        local_kw = {"wav": mine}
        known_all = not local_input
    if pad_to is None:
        if known_all:
            pad_to = max(int(torch.as_tensor(w).shape[-1]) for w in items)
        else:
            local_max = max((int(torch.as_tensor(w).shape[-1]) for w in local_kw["wav"]), default=0)
            pad_to = global_max_length(local_max, device=dev, group=group)
    per_rank = int(per_rank) if sizes_in_block else max(sizes)
    all_feat_h = None
    if hasattr(segmenter, "call_with_tables") and not gather_features:
        # fast path: the device-side segment table goes into the collective as it is - ONE all-gather of a
        # (per_rank, 1 + T, 2) int32 block whose row 0 carries the count, then one device->host copy
        import time as _time
        _t = [_time.perf_counter()]
        eng = segmenter._engine
        T = eng.num_frames(pad_to)
        gdev = eng.device
        block = torch.zeros((per_rank, 1 + T, 2), dtype=torch.int32, device=gdev)
        comm = getattr(eng, "_comm_stream", None)
        if comm is None:
            comm = eng._comm_stream = torch.cuda.Stream(device=gdev)
        state = {}

        def exchange(streams, seg_dev, cnt_dev):
            # enqueued behind the sub-batch streams while the host still waits for the last hidden states: fill the block,
            # ONE all-gather, one copy of the gathered table into pinned memory
            issuing = torch.cuda.current_stream(gdev)
            with torch.cuda.stream(comm):
                comm.wait_stream(issuing)                 # `block` was zeroed on the caller's stream
                for st in streams:
                    comm.wait_stream(st)
                if seg_dev is not None:
                    block[:seg_dev.shape[0], 1:] = seg_dev
                    block[:seg_dev.shape[0], 0, 0] = cnt_dev
                    block[0, 0, 1] = seg_dev.shape[0]     # header: how many utterances this rank holds
                src = block.to(dev)                       # gloo: host tensors (synchronises); NCCL: the device block itself
                gathered = torch.empty((world * per_rank, 1 + T, 2), dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(gathered, src, group=group)
                if gathered.is_cuda:
                    pin = torch.empty(gathered.shape, dtype=torch.int32, pin_memory=True)
                    pin.copy_(gathered, non_blocking=True)
                    state["ev"] = torch.cuda.Event()
                    state["ev"].record(comm)
                    state["g"] = pin
                else:
                    state["g"] = gathered

        _t.append(_time.perf_counter())
        if mine:
            local, _, _ = segmenter.call_with_tables(local_kw["wav"], pad_to=pad_to, after_enqueue=exchange, hidden_to=hidden_to)
        else:
            local = []
            exchange([], None, None)
        _t.append(_time.perf_counter())
        if "ev" in state:
            state["ev"].synchronize()
        _t.append(_time.perf_counter())
        g = state["g"].numpy()
        all_cnt_h, all_seg_h = g[:, 0, 0], g[:, 1:]
        if sizes_in_block:
            sizes = [int(g[r * per_rank, 0, 1]) for r in range(world)]
    else:
        local = segmenter(in_second=False, pad_to=pad_to, **({"hidden_to": hidden_to} if hidden_to != "host" else {}), **local_kw) if mine else []
        # (1) counts, (2) fixed-stride table
        cnt = torch.zeros(per_rank, dtype=torch.int32)
        for k, r in enumerate(local):
            cnt[k] = len(r["segments"])
        all_cnt = torch.empty(world * per_rank, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_cnt, cnt.to(dev), group=group)
        all_cnt_h = all_cnt.cpu().numpy()
        stride = max(int(all_cnt_h.max()), 1)
        seg = torch.zeros((per_rank, stride, 2), dtype=torch.int32)
        for k, r in enumerate(local):
            n = int(cnt[k])
            if n:
                seg[k, :n] = torch.from_numpy(np.asarray(r["segments"], dtype=np.int32).reshape(-1, 2))
        all_seg = torch.empty((world * per_rank, stride, 2), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_seg, seg.to(dev), group=group)
        all_seg_h = all_seg.cpu().numpy()
        if gather_features:
            feat = torch.zeros((per_rank, stride, 768), dtype=torch.float32)
            for k, r in enumerate(local):
                n = int(cnt[k])
                if n:
                    feat[k, :n] = torch.from_numpy(np.ascontiguousarray(r["segment_features"]))
            all_feat = torch.empty((world * per_rank, stride, 768), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(all_feat, feat.to(dev), group=group)
            all_feat_h = all_feat.cpu().numpy()
    # unpack: ONE int64 (and one seconds) conversion of the occupied part of the table, then per-utterance views - the loop
    # runs over every utterance of the WHOLE batch on every rank, so it must not allocate or convert per row
    counts = np.asarray(all_cnt_h).tolist()
    n_max = max(max(counts), 1)
    seg64 = np.ascontiguousarray(all_seg_h[:, :n_max], dtype=np.int64)
    table = seg64 * 1.0 / 50 if in_second else seg64       # the reference's `segments * 1.0 / 50` (sylber.py:132)
    empty = np.array([]) * 1.0 / 50 if in_second else np.array([])
    out = []
    for r in range(world):
        base = r * per_rank
        mine_r = r == rank
        for k in range(sizes[r]):
            n = counts[base + k]
            d = {"segments": table[base + k, :n] if n > 0 else empty, "segment_features": None, "hidden_states": None}
            if mine_r:
                d["segment_features"] = local[k]["segment_features"]
                d["hidden_states"] = local[k]["hidden_states"]
            elif all_feat_h is not None:
                d["segment_features"] = all_feat_h[base + k, :n] if n > 0 else np.array([])
            out.append(d)
    if "_t" in locals():
        _t.append(_time.perf_counter())
        LAST_TIMING.update(setup=_t[1] - _t[0], call=_t[2] - _t[1], wait_gather=_t[3] - _t[2], unpack=_t[4] - _t[3])
    return out
