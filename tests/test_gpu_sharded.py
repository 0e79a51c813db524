"""The product-level sharded call (sylber_b200.distributed.segment_sharded) on real GPUs: the result of a list sharded
over `world` ranks is BIT-IDENTICAL to one process calling `Segmenter(wav=list)` - uniform and mixed lengths
(SURVEY.md 4 "sharded vs single-GPU result, bit-identical"; 8e: every rank pads to the global T_max).

With >= world GPUs every rank owns one GPU and the segment table travels over NCCL; on a one-GPU box the ranks share
cuda:0 and the table travels over gloo (NCCL refuses two ranks on one device) - the forward, the padding rule and the
table packing are the same code either way."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _clips(kind):
    g = torch.Generator().manual_seed(21)
    if kind == "uniform":
        return [torch.randn(1, 48000, generator=g) for _ in range(6)]
    lens = [48000, 16000, 31234, 9000, 40000, 22222, 12000]
    return [torch.randn(1, n, generator=g) for n in lens]


def _worker(rank, world, port, kind, nccl, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if nccl else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", dev)} if nccl else {}))
    try:
        from sylber_b200 import Segmenter, segment_sharded
        from sylber_b200.distributed import shard_range
        from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
        seg = Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM), device=f"cuda:{dev}")
        clips = _clips(kind)
        res_global = segment_sharded(seg, wav=clips, in_second=False, gather_features=True)
        lo, hi = shard_range(len(clips), rank, world)
        res_local = segment_sharded(seg, wav=clips[lo:hi], in_second=True, local_input=True)
        # with the caller's bound on the shard size the list lengths travel inside the table block (no size collective)
        res_hint = segment_sharded(seg, wav=clips[lo:hi], in_second=True, local_input=True, per_rank=(len(clips) + world - 1) // world)
        assert len(res_hint) == len(res_local) and all(np.array_equal(np.asarray(a["segments"]), np.asarray(b["segments"]))
                                                       for a, b in zip(res_hint, res_local))
        q.put((rank, [(np.asarray(r["segments"]), r["segment_features"], r["hidden_states"]) for r in res_global],
               [np.asarray(r["segments"]) for r in res_local]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["uniform", "mixed"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_call_is_bit_identical_to_the_single_call(cuda, world, kind):
    nccl = torch.cuda.device_count() >= world
    if world == 4 and not nccl:
        pytest.skip("world 4 runs with one GPU per rank only")
    from sylber_b200 import Segmenter
    from sylber_b200.distributed import shard_range
    from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
    clips = _clips(kind)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, nccl, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, res_global, res_local = q.get(timeout=600)
        results[rank] = (res_global, res_local)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM), device="cuda:0")(
        wav=clips, in_second=False)
    assert sum(len(o["segments"]) for o in single) > 20          # the comparison is not vacuous
    for rank, (res_global, res_local) in results.items():
        lo, hi = shard_range(len(clips), rank, world)
        assert len(res_global) == len(res_local) == len(clips)
        for i, (seg, feat, hid) in enumerate(res_global):
            want = np.asarray(single[i]["segments"])
            assert seg.shape == want.shape and np.array_equal(seg, want), (rank, i)
            assert np.array_equal(np.asarray(feat), np.asarray(single[i]["segment_features"])), (rank, i)   # gathered features
            if lo <= i < hi:
                assert np.array_equal(hid, single[i]["hidden_states"]), (rank, i)     # padded to the GLOBAL maximum
            else:
                assert hid is None
            assert np.array_equal(res_local[i], want * 1.0 / 50 if len(want) else want)   # per-rank input lists, seconds


def test_two_devices_in_one_process(cuda):
    """the >48 KB shared-memory opt-in is per device, and entry points must not move the caller's current
    device.  Two Segmenters on two GPUs of ONE process give identical results and leave torch's current device alone."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from sylber_b200 import Segmenter
    from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
    sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
    clips = _clips("mixed")[:3]
    assert torch.cuda.current_device() == 0
    a = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0")
    b = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:1")      # second handle, other device: needs its own attributes
    assert torch.cuda.current_device() == 0
    ra = a(wav=clips, in_second=False)
    rb = b(wav=clips, in_second=False)
    assert torch.cuda.current_device() == 0
    ra2 = a(wav=clips, in_second=False)                                  # and back on device 0 after device 1 was used
    for x, y, z in zip(ra, rb, ra2):
        assert np.array_equal(x["hidden_states"], y["hidden_states"]) and np.array_equal(x["hidden_states"], z["hidden_states"])
        assert np.array_equal(np.asarray(x["segments"]), np.asarray(y["segments"]))
