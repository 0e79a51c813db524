"""Checkpoint handling for the Segmenter: key normalisation and synthetic (random) weights.

The reference loads a bare `HubertModel` state_dict with strict=False (sylber/model/sylber.py:51-52), which
silently ignores missing or renamed tensors.  Here every tensor the forward path reads is required.
Accepted containers: a bare state_dict, a Lightning checkpoint (`['state_dict']`, keys prefixed
`net.speech_model.` - sylber_trainer.py:21, train.py:14), and the EMA dump (`['ema']`, train.py:17,24).
"""
from __future__ import annotations

import math

import torch

_POS = "encoder.pos_conv_embed.conv."
_PREFIXES = ("net.speech_model.", "model.speech_model.", "speech_model.", "net.", "module.")


def REQUIRED_KEYS(n_layers):
    keys = ["feature_extractor.conv_layers.0.conv.weight",
            "feature_extractor.conv_layers.0.layer_norm.weight",
            "feature_extractor.conv_layers.0.layer_norm.bias"]
    keys += [f"feature_extractor.conv_layers.{i}.conv.weight" for i in range(1, 7)]
    keys += ["feature_projection.layer_norm.weight", "feature_projection.layer_norm.bias",
             "feature_projection.projection.weight", "feature_projection.projection.bias",
             _POS + "bias", _POS + "parametrizations.weight.original0", _POS + "parametrizations.weight.original1",
             "encoder.layer_norm.weight", "encoder.layer_norm.bias"]
    for l in range(n_layers):
        p = f"encoder.layers.{l}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            keys += [p + f"attention.{proj}.weight", p + f"attention.{proj}.bias"]
        keys += [p + "layer_norm.weight", p + "layer_norm.bias",
                 p + "feed_forward.intermediate_dense.weight", p + "feed_forward.intermediate_dense.bias",
                 p + "feed_forward.output_dense.weight", p + "feed_forward.output_dense.bias",
                 p + "final_layer_norm.weight", p + "final_layer_norm.bias"]
    return keys


def normalize_state_dict(obj):
    """Return a flat {HubertModel key: tensor} dict from any of the accepted checkpoint containers."""
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    elif isinstance(obj, dict) and "ema" in obj and isinstance(obj["ema"], dict):
        obj = obj["ema"]
    if hasattr(obj, "state_dict") and not isinstance(obj, dict):
        obj = obj.state_dict()
    out = {}
    for key, val in obj.items():
        if not torch.is_tensor(val):
            continue
        for pre in _PREFIXES:
            if key.startswith(pre) and (key[len(pre):].startswith(("feature_", "encoder.", "masked_spec"))):
                key = key[len(pre):]
                break
        if key == _POS + "weight_g":
            key = _POS + "parametrizations.weight.original0"
        elif key == _POS + "weight_v":
            key = _POS + "parametrizations.weight.original1"
        out[key] = val
    if _POS + "weight" in out and _POS + "parametrizations.weight.original1" not in out:
        # already-folded positional conv weight: express it as g = per-tap norm, v = weight
        w = out.pop(_POS + "weight").float()
        out[_POS + "parametrizations.weight.original1"] = w
        out[_POS + "parametrizations.weight.original0"] = w.norm(2, dim=(0, 1), keepdim=True)
    return out


def random_hubert_state_dict(n_layers=9, seed=0):
    """Random weights with the hubert-base architecture and HF-like initial scales (synthetic data for
    tests and benchmarks - there is no checkpoint offline)."""
    g = torch.Generator().manual_seed(seed)

    def normal(*shape, std):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    fe = "feature_extractor.conv_layers."
    kernels = (10, 3, 3, 3, 3, 2, 2)
    for i, k in enumerate(kernels):
        cin = 1 if i == 0 else 512
        sd[fe + f"{i}.conv.weight"] = normal(512, cin, k, std=math.sqrt(2.0 / (cin * k)))   # kaiming normal
    sd[fe + "0.layer_norm.weight"] = torch.ones(512)
    sd[fe + "0.layer_norm.bias"] = torch.zeros(512)
    sd["feature_projection.layer_norm.weight"] = torch.ones(512)
    sd["feature_projection.layer_norm.bias"] = torch.zeros(512)
    k = math.sqrt(1.0 / 512)
    sd["feature_projection.projection.weight"] = (torch.rand(768, 512, generator=g) * 2 - 1) * k
    sd["feature_projection.projection.bias"] = (torch.rand(768, generator=g) * 2 - 1) * k
    v = normal(768, 48, 128, std=2 * math.sqrt(1.0 / (128 * 768)))
    sd[_POS + "parametrizations.weight.original1"] = v
    sd[_POS + "parametrizations.weight.original0"] = v.norm(2, dim=(0, 1), keepdim=True)
    sd[_POS + "bias"] = torch.zeros(768)
    sd["encoder.layer_norm.weight"] = torch.ones(768)
    sd["encoder.layer_norm.bias"] = torch.zeros(768)
    for l in range(n_layers):
        p = f"encoder.layers.{l}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"attention.{proj}.weight"] = normal(768, 768, std=0.02)
            sd[p + f"attention.{proj}.bias"] = torch.zeros(768)
        sd[p + "feed_forward.intermediate_dense.weight"] = normal(3072, 768, std=0.02)
        sd[p + "feed_forward.intermediate_dense.bias"] = torch.zeros(3072)
        sd[p + "feed_forward.output_dense.weight"] = normal(768, 3072, std=0.02)
        sd[p + "feed_forward.output_dense.bias"] = torch.zeros(768)
        for ln in ("layer_norm", "final_layer_norm"):
            sd[p + ln + ".weight"] = torch.ones(768)
            sd[p + ln + ".bias"] = torch.zeros(768)
    return sd


SPEECH_LIKE_BIAS_NORM = 2.2


def syllabic_test_state_dict(n_layers=9, seed=0, bias_norm=1.9):
    """Random weights whose last LayerNorm is rescaled so that the segmentation exercises every branch
    (norm mask on/off, merges, splits, refinements) instead of the degenerate 'every frame is a segment'
    behaviour of plain random weights.  Biases are non-zero so bias paths are covered.

    `bias_norm` places the frame norms relative to the 2.6 threshold (measured on 10 s of N(0,1) audio, seed 0):
    1.9 (the value the golden files were generated with) -> median norm 2.43, 3-5 % of the frames above the threshold,
    ~20 short segments; 2.1 -> median 2.60, half the frames on, ~125 segments; SPEECH_LIKE_BIAS_NORM = 2.2 -> median
    2.68, 85 % of the frames on, ~65 segments of up to 30 frames, ~360 merge decisions per utterance - the occupancy
    of real speech (README.md:5: 4.27 segments per second), used by bench.py and the agreement tests."""
    sd = random_hubert_state_dict(n_layers, seed)
    g = torch.Generator().manual_seed(seed + 3)
    for key in list(sd):
        if key.endswith(".bias") and "layer_norm" not in key:
            sd[key] = torch.randn(sd[key].shape, generator=g) * 0.05
        elif key.endswith("layer_norm.weight"):
            sd[key] = 1.0 + 0.1 * torch.randn(sd[key].shape, generator=g)
        elif key.endswith("layer_norm.bias"):
            sd[key] = 0.1 * torch.randn(sd[key].shape, generator=g)
    last = f"encoder.layers.{n_layers - 1}.final_layer_norm."
    sd[last + "weight"] = (torch.rand(768, generator=g) < 0.05).float() * 0.28
    b = torch.randn(768, generator=g)
    sd[last + "bias"] = b / b.norm() * float(bias_norm)
    return sd
