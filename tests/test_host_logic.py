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
