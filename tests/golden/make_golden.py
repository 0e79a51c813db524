"""Generates tests/golden/*.npz by running the UNMODIFIED reference from /root/reference (SURVEY.md 8c recipe).

The reference ships no tests and no golden vectors, so the known answers are produced here, once, in the build
container (the GPU box has no /root/reference), and committed together with this script:
  segment_cases.npz     reference get_segment + segment means on synthetic syllable-like states
  sample_wav.npz        reference Segmenter(...)(wav=...) on samples/sample.wav with seeded synthetic weights
                        (no checkpoint exists offline): segments, segment_features, hidden_states
  hubert_padded.npz     transformers.HubertModel last_hidden_state on a padded 2-clip batch
Every file records torch / transformers / numpy versions and the seeds.

Run:  python tests/golden/make_golden.py          (needs /root/reference; refuses to run without it)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def import_reference():
    """Stub the five uninstalled modules the hot path never touches, then import the reference package."""
    if not os.path.isdir(REF):
        raise SystemExit("make_golden.py needs the reference checkout at /root/reference")
    stubs = {"torchode": ["Tsit5"], "torchdiffeq": ["odeint"], "gateloop_transformer": ["SimpleGateLoopLayer"],
             "vector_quantize_pytorch": ["GroupedResidualVQ"], "lightning": []}
    for name, attrs in stubs.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    sys.modules["lightning"].LightningModule = torch.nn.Module
    sys.path.insert(0, REF)
    import sylber  # noqa: F401
    from sylber.utils.segment_utils import get_segment
    return sylber, get_segment


def versions():
    import transformers
    return np.array([f"torch={torch.__version__}", f"transformers={transformers.__version__}", f"numpy={np.__version__}"])


def main():
    from seg_cases import plateau_states
    from sylber_b200.weights import syllabic_test_state_dict
    from transformers import HubertConfig, HubertModel
    sylber, get_segment = import_reference()

    # ---- 1. segmentation cases
    rng = np.random.default_rng(20261017)
    out = {"versions": versions(), "seed": np.array([20261017])}
    lengths = [1, 2, 7, 16, 33, 48, 64, 64]
    for i, T in enumerate(lengths):
        st = plateau_states(rng, T)
        if i == 2:
            st[:] *= 0.01                       # everything below the norm threshold -> empty result
        seg = get_segment(st, 2.6, 0.8)
        out[f"states_{i}"] = st
        out[f"segments_{i}"] = np.asarray(seg)
        out[f"features_{i}"] = (np.stack([st[s:e].mean(0) for s, e in seg]) if len(seg) > 0 else np.array([]))
    np.savez_compressed(os.path.join(HERE, "segment_cases.npz"), **out)
    print("segment_cases:", [out[f"segments_{i}"].shape for i in range(len(lengths))])

    # ---- 2. reference Segmenter on samples/sample.wav with the synthetic 'syllabic' weights
    from scipy.io import wavfile
    sr, data = wavfile.read(os.path.join(REF, "samples", "sample.wav"))
    assert sr == 16000 and data.dtype == np.int16
    wav = torch.from_numpy(data.astype(np.float32) / 32768.0)[None, :]
    wav = (wav - wav.mean()) / wav.std()                      # sylber/model/sylber.py:86
    sd = syllabic_test_state_dict(9, seed=0)
    with tempfile.TemporaryDirectory() as d:
        HubertConfig().save_pretrained(d)
        seg = sylber.Segmenter(model_ckpt=None, speech_upstream=d, device="cpu")
    missing, unexpected = seg.speech_model.load_state_dict(sd, strict=False)
    assert missing == ["masked_spec_embed"] and not unexpected, (missing, unexpected)
    res = seg(wav=wav, in_second=False)
    res_s = seg(wav=wav, in_second=True)
    # a padded two-clip batch through the same reference call (list input, mixed lengths)
    g = torch.Generator().manual_seed(5)
    w2 = [wav[:, :30000].clone(), torch.randn(1, 17000, generator=g)]
    res2 = seg(wav=w2, in_second=False)
    np.savez_compressed(
        os.path.join(HERE, "sample_wav.npz"), versions=versions(), weights_seed=np.array([0]), wav=wav.numpy(),
        segments=res["segments"], segments_sec=res_s["segments"], segment_features=res["segment_features"],
        hidden_states=res["hidden_states"],
        b_wav0=w2[0].numpy(), b_wav1=w2[1].numpy(),
        b_segments0=res2[0]["segments"], b_segments1=res2[1]["segments"],
        b_features0=res2[0]["segment_features"], b_features1=res2[1]["segment_features"],
        b_hidden0=res2[0]["hidden_states"], b_hidden1=res2[1]["hidden_states"])
    print("sample_wav: segments", res["segments"].shape, "hidden", res["hidden_states"].shape,
          "| padded batch:", res2[0]["segments"].shape, res2[1]["segments"].shape)

    # ---- 3. HubertModel on a padded batch (pins oracle/hubert_ref.py against the third-party dependency)
    torch.manual_seed(0)
    model = HubertModel(HubertConfig(num_hidden_layers=3)).eval()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 16000, generator=g)
    x[1, 9000:] = 0
    mask = torch.ones(2, 16000, dtype=torch.long)
    mask[1, 9000:] = 0
    with torch.no_grad():
        hs = model(x, attention_mask=mask).last_hidden_state
    np.savez_compressed(os.path.join(HERE, "hubert_padded.npz"), versions=versions(), init_seed=np.array([0]),
                        input_seed=np.array([7]), n_layers=np.array([3]), lengths=np.array([16000, 9000]),
                        hidden=hs.numpy())
    print("hubert_padded:", tuple(hs.shape))


if __name__ == "__main__":
    main()
