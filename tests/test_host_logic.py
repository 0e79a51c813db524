"""Host-side logic that needs no GPU: checkpoint normalisation, key inventory, sharding arithmetic."""
import numpy as np
import pytest
import torch

from sylber_b200.weights import (REQUIRED_KEYS, normalize_state_dict, random_hubert_state_dict,
                                 syllabic_test_state_dict)
from sylber_b200.distributed import shard_range

POS = "encoder.pos_conv_embed.conv."


def test_required_keys_match_hubert_state_dict():
    from transformers import HubertConfig, HubertModel
    model = HubertModel(HubertConfig(num_hidden_layers=9))
    sd = model.state_dict()
    assert len(sd) == 163
    req = REQUIRED_KEYS(9)
    assert set(req) == set(sd) - {"masked_spec_embed"}
    ours = random_hubert_state_dict(9)
    for k in req:
        assert tuple(ours[k].shape) == tuple(sd[k].shape), k


def test_normalize_accepts_all_containers():
    sd = random_hubert_state_dict(2)
    req = REQUIRED_KEYS(2)
    assert set(req) <= set(normalize_state_dict(sd))
    lightning = {"state_dict": {"net.speech_model." + k: v for k, v in sd.items()}, "epoch": 3}
    assert set(req) <= set(normalize_state_dict(lightning))
    ema = {"ema": {"speech_model." + k: v for k, v in sd.items()}}
    assert set(req) <= set(normalize_state_dict(ema))
    legacy = dict(sd)
    legacy[POS + "weight_g"] = legacy.pop(POS + "parametrizations.weight.original0")
    legacy[POS + "weight_v"] = legacy.pop(POS + "parametrizations.weight.original1")
    out = normalize_state_dict(legacy)
    assert set(req) <= set(out) and POS + "weight_g" not in out
    folded = {k: v for k, v in sd.items() if "parametrizations" not in k}
    w = torch.randn(768, 48, 128)
    folded[POS + "weight"] = w
    out = normalize_state_dict(folded)
    g, v = out[POS + "parametrizations.weight.original0"], out[POS + "parametrizations.weight.original1"]
    assert torch.allclose(v * (g / v.norm(2, dim=(0, 1), keepdim=True)), w, atol=1e-6)


def test_syllabic_weights_are_deterministic():
    a, b = syllabic_test_state_dict(9, 0), syllabic_test_state_dict(9, 0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    last = "encoder.layers.8.final_layer_norm."
    assert float(a[last + "bias"].norm()) == pytest.approx(1.9, rel=1e-5)


@pytest.mark.parametrize("n,world", [(256, 8), (10, 4), (3, 8), (0, 2), (33, 2)])
def test_shard_range_partitions(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def _cuh_floats(src, start, end):
    """float literals (with an f suffix) between two markers of a .cuh source"""
    import re
    seg = src[src.index(start):src.index(end, src.index(start))]
    return [float(x[:-1]) for x in re.findall(r"(-?\d+\.\d+(?:e-?\d+)?f)", seg)]


def test_device_gelu_formula_error_bound():
    """gelu_fast2 (csrc/common.cuh): max(x,0) - |x| 2^L(min(|x|,6)), L a degree-6 fit.  The constants are read from the
    source and the formula is replayed in float32; the bound quoted there (2.9e-7 absolute) is re-derived."""
    import os
    from scipy.special import erf
    src = open(os.path.join(os.path.dirname(__file__), "..", "sylber_b200", "csrc", "common.cuh")).read()
    c = _cuh_floats(src, "__device__ __forceinline__ void gelu_fast2(", "float a0, a1, e0, e1;")
    coeffs = [v for i, v in enumerate(c) if v not in (6.0,)][0::2]          # every constant appears twice (pack2)
    assert len(coeffs) == 7, coeffs
    f = np.float32
    x = np.concatenate([np.linspace(-20, 20, 2000001), np.random.default_rng(0).normal(size=500000) * 1.5]).astype(f)
    t = np.minimum(np.abs(x), f(6.0)).astype(f)
    q = np.full_like(t, f(coeffs[0]))
    for k in coeffs[1:]:
        q = (q * t + f(k)).astype(f)
    g = (np.maximum(x, f(0)) - np.abs(x) * np.exp2(q).astype(f)).astype(f)
    xd = x.astype(np.float64)
    ref = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
    err = np.abs(g - ref)
    assert err.max() < 3.5e-7, err.max()
    assert err[np.abs(x) < 1].max() < 1.6e-7


def test_device_exp2_polynomial_error_bound():
    """exp2_poly2 (csrc/attention.cuh): Cody-Waite split + degree-4 polynomial, relative error < 3e-6."""
    import os
    src = open(os.path.join(os.path.dirname(__file__), "..", "sylber_b200", "csrc", "attention.cuh")).read()
    c = _cuh_floats(src, "f32x2 q = fma2(pack2(0.0095", "float t0, t1, q0, q1;")
    coeffs = c[0::2]
    assert len(coeffs) == 5, coeffs
    f = np.float32
    a = np.linspace(-30, 8, 1000001).astype(f)
    n = np.rint(a).astype(f)
    fr = (a - n).astype(f)
    q = np.full_like(fr, f(coeffs[0]))
    for k in coeffs[1:]:
        q = (q * fr + f(k)).astype(f)
    got = np.ldexp(q.astype(np.float64), n.astype(np.int64))
    rel = np.abs(got / np.exp2(a.astype(np.float64)) - 1)
    assert rel.max() < 3.5e-6, rel.max()


def _streamk_items(cluster, clusters, tiles, kb_total):
    """The walk of gemm3_tc_kernel<true> (csrc/gemm3_tc.cuh, next_item): cluster `cluster` owns the k-block units
    [c U / C, (c + 1) U / C) and visits them from the end of the range to its start, one (tile, kb0, kb1) item at a
    time."""
    units = tiles * kb_total
    lo, pos = units * cluster // clusters, units * (cluster + 1) // clusters
    items = []
    while pos > lo:
        tile = (pos - 1) // kb_total
        t0 = tile * kb_total
        start = max(t0, lo)
        items.append((tile, start - t0, pos - t0))
        pos = start
    return items


@pytest.mark.parametrize("tiles,kb_total", [(189, 12), (189, 48), (192, 8), (768, 12), (256, 48), (128, 48), (576, 12),
                                            (75, 1), (147, 3), (4032, 24), (1000, 7)])
def test_streamk_schedule_invariants(tiles, kb_total):
    """What the kernel's fix-up protocol relies on, checked on the schedule arithmetic for the GEMM shapes of the
    batch 32 x 10 s forward and a few odd ones (the host only selects the schedule when tiles > clusters):
    every k-block of every tile is computed exactly once; a tile is cut at most once; a cluster's cut items are its
    FIRST (a head: k-blocks [0, kb1), parked for the next cluster) and its LAST (a tail: k-blocks [kb0, end), which
    adds the previous cluster's partial) and nothing in between; the tail of cluster c is the tile whose head
    cluster c - 1 parked."""
    clusters = 74
    assert tiles > clusters
    seen = np.zeros((tiles, kb_total), dtype=np.int32)
    heads, tails = {}, {}
    for c in range(clusters):
        items = _streamk_items(c, clusters, tiles, kb_total)
        assert items, "every cluster has work"
        for i, (tile, kb0, kb1) in enumerate(items):
            assert 0 <= kb0 < kb1 <= kb_total
            seen[tile, kb0:kb1] += 1
            assert not (kb0 > 0 and kb1 < kb_total), "an item is never cut on both sides"
            if kb1 < kb_total:
                assert i == 0, "a head item is the first thing a cluster does"
                heads[c] = tile
            if kb0 > 0:
                assert i == len(items) - 1, "a tail item is the last thing a cluster does"
                tails[c] = tile
    assert (seen == 1).all()
    assert 0 not in tails and clusters - 1 not in heads
    assert {c + 1: t for c, t in heads.items()} == tails
