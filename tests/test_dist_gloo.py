"""World-size-2 gloo test of the only exchange on the path: agreeing on T_max and all-gathering the segment table."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sylber_b200.distributed import (gather_segment_table, global_max_length, segment_sharded, shard_range,
                                     unpack_segment_table)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lengths = [16000 + 977 * i for i in range(n_items)]
        lo, hi = shard_range(n_items, rank, world)
        t_max = global_max_length(max(lengths[lo:hi]) if hi > lo else 0)
        per_rank = -(-n_items // world)
        max_seg = 6
        # fake "local result": utterance g has (g % 5) segments [10*g + k, 10*g + k + 1)
        seg = torch.zeros((per_rank, max_seg, 2), dtype=torch.int32)
        cnt = torch.zeros((per_rank,), dtype=torch.int32)
        for k, gidx in enumerate(range(lo, hi)):
            n = gidx % 5
            cnt[k] = n
            for j in range(n):
                seg[k, j, 0] = 10 * gidx + j
                seg[k, j, 1] = 10 * gidx + j + 1
        all_seg, all_cnt = gather_segment_table(seg, cnt)
        table = unpack_segment_table(all_seg, all_cnt, n_items, world)
        q.put((rank, t_max, [t.tolist() for t in table]))
    finally:
        dist.destroy_process_group()


def test_gather_segment_table_world2():
    world, n_items = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[[10 * g + j, 10 * g + j + 1] for j in range(g % 5)] for g in range(n_items)]
    for rank, t_max, table in results:
        assert t_max == 16000 + 977 * (n_items - 1)           # every rank pads to the global maximum
        assert table == want                                   # and sees the whole batch's boundaries, in order


def test_single_process_passthrough():
    seg = torch.arange(12, dtype=torch.int32).view(2, 3, 2)
    cnt = torch.tensor([3, 1], dtype=torch.int32)
    a, b = gather_segment_table(seg, cnt)
    assert a is seg and b is cnt
    assert global_max_length(123) == 123
    t = unpack_segment_table(seg, cnt, 2, 1)
    assert t[0].shape == (3, 2) and t[1].tolist() == [[6, 7]]


class _FakeSegmenter:
    """Stands in for the CUDA Segmenter on a CPU box: a deterministic function of (clip, padded length), which is exactly
    what the real one is, so a wrong shard range, order or pad_to shows up in the comparison."""
    device = "cpu"

    def __call__(self, wav_file=None, wav=None, in_second=True, pad_to=None):
        clips = wav if isinstance(wav, list) else [wav]
        t_max = max(max(int(w.shape[-1]) for w in clips), int(pad_to or 0))
        out = []
        for w in clips:
            n = int(w.shape[-1])
            k = int(abs(float(w[0, 0])) * 10) % 4                     # 0..3 segments, 0 exercises the empty contract
            seg = np.array([[j * 7 + n % 5, j * 7 + 3 + t_max % 3] for j in range(k)], dtype=np.int64) if k else np.array([])
            out.append({"segments": seg * 1.0 / 50 if in_second else seg,
                        "segment_features": np.full((k, 768), float(w[0, 0]), np.float32) if k else np.array([]),
                        "hidden_states": np.full((t_max // 320, 768), float(w[0, -1]), np.float32)})
        return out if isinstance(wav, list) else out[0]


def _clips(n_items):
    g = torch.Generator().manual_seed(11)
    return [torch.randn(1, 4000 + 613 * ((i * 5) % n_items), generator=g) for i in range(n_items)]


def _sharded_worker(rank, world, port, n_items, local_input, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clips = _clips(n_items)
        if local_input:
            lo, hi = shard_range(n_items, rank, world)
            res = segment_sharded(_FakeSegmenter(), wav=clips[lo:hi], in_second=False, local_input=True, gather_features=True)
        else:
            res = segment_sharded(_FakeSegmenter(), wav=clips, in_second=False, gather_features=(rank == 0) or True)
        q.put((rank, [(np.asarray(r["segments"]).tolist(), None if r["segment_features"] is None else float(np.asarray(r["segment_features"]).sum()),
                       None if r["hidden_states"] is None else r["hidden_states"].shape[0]) for r in res]))
    finally:
        dist.destroy_process_group()


def _run_sharded(world, n_items, local_input):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n_items, local_input, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_segment_sharded_matches_single_call_world2():
    """The product-level sharded call (global list and per-rank lists) against one un-sharded call: same segments in the
    same order on every rank, features gathered, hidden states only for the rank's own block and of the GLOBAL T_max."""
    for n_items, local_input in ((7, False), (7, True), (1, False)):       # 1 item: rank 1's shard is empty
        single = _FakeSegmenter()(wav=_clips(n_items), in_second=False)
        results = _run_sharded(2, n_items, local_input)
        for rank, res in results.items():
            lo, hi = shard_range(n_items, rank, 2)
            assert len(res) == n_items
            for i, (seg, feat_sum, t_rows) in enumerate(res):
                assert seg == np.asarray(single[i]["segments"]).tolist(), (n_items, local_input, rank, i)
                assert feat_sum == float(np.asarray(single[i]["segment_features"]).sum())
                if lo <= i < hi:
                    assert t_rows == single[i]["hidden_states"].shape[0]   # padded to the global maximum
                else:
                    assert t_rows is None
