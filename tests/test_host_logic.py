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
