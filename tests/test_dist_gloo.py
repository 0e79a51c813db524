"""World-size-2 gloo test of the only exchange on the path: agreeing on T_max and all-gathering the segment table."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sylber_b200.distributed import gather_segment_table, global_max_length, shard_range, unpack_segment_table


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
