"""oracle/hubert_ref.py (fp32 restatement of the HuBERT forward) against the golden vectors produced by
transformers.HubertModel / the reference Segmenter, and against the live third-party model when importable."""
import os

import numpy as np
import pytest
import torch

from oracle.hubert_ref import hubert_forward, num_frames, conv_out_lengths
from oracle import segment_ref as R
from sylber_b200.weights import syllabic_test_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_frame_counts():
    assert num_frames(46080) == 143          # samples/sample.wav
    assert num_frames(160000) == 499
    assert num_frames(960000) == 2999
    assert conv_out_lengths(160000) == [31999, 15999, 7999, 3999, 1999, 999, 499]
    assert num_frames(400) == 1 and num_frames(399) == 0


def test_sample_wav_golden():
    g = np.load(os.path.join(GOLD, "sample_wav.npz"))
    sd = syllabic_test_state_dict(9, seed=int(g["weights_seed"][0]))
    hs = hubert_forward(sd, torch.from_numpy(g["wav"]), [g["wav"].shape[1]], 9).numpy()[0]
    assert hs.shape == (143, 768)
    assert _rel(hs, g["hidden_states"]) < 2e-5
    # segmentation of the reference's own states reproduces the reference's segments exactly
    seg = R.get_segment(g["hidden_states"], 2.6, 0.8)
    assert np.array_equal(seg, g["segments"])
    assert np.array_equal(R.c_get_segment(g["hidden_states"], 2.6, 0.8), g["segments"])
    assert np.array_equal(seg * 1.0 / 50, g["segments_sec"])
    assert np.array_equal(R.c_segment_mean(g["hidden_states"], seg), g["segment_features"])
    # coverage of the branches (H5): frames on both sides of the norm threshold, merges and splits
    norms = np.sqrt((g["hidden_states"] ** 2).sum(-1))
    assert 0.05 < (norms >= 2.6).mean() < 0.999
    assert 1 < len(seg) < 143


def test_padded_batch_golden():
    g = np.load(os.path.join(GOLD, "sample_wav.npz"))
    sd = syllabic_test_state_dict(9, seed=0)
    w0, w1 = torch.from_numpy(g["b_wav0"]), torch.from_numpy(g["b_wav1"])
    n = max(w0.shape[1], w1.shape[1])
    batch = torch.zeros(2, n)
    batch[0, :w0.shape[1]] = w0[0]
    batch[1, :w1.shape[1]] = w1[0]
    hs = hubert_forward(sd, batch, [w0.shape[1], w1.shape[1]], 9).numpy()
    assert _rel(hs[0], g["b_hidden0"]) < 2e-5 and _rel(hs[1], g["b_hidden1"]) < 2e-5
    # padded frames are returned and segmented: hidden has T_max rows for the short clip too
    assert g["b_hidden1"].shape[0] == num_frames(n)


def test_hubert_padded_golden_and_live():
    g = np.load(os.path.join(GOLD, "hubert_padded.npz"))
    from transformers import HubertConfig, HubertModel
    torch.manual_seed(int(g["init_seed"][0]))
    model = HubertModel(HubertConfig(num_hidden_layers=int(g["n_layers"][0]))).eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(int(g["input_seed"][0]))
    x = torch.randn(2, 16000, generator=gen)
    x[1, 9000:] = 0
    hs = hubert_forward(sd, x, g["lengths"].tolist(), int(g["n_layers"][0])).numpy()
    assert _rel(hs, g["hidden"]) < 2e-5
    mask = torch.ones(2, 16000, dtype=torch.long)
    mask[1, 9000:] = 0
    with torch.no_grad():
        live = model(x, attention_mask=mask).last_hidden_state.numpy()
    assert _rel(hs, live) < 2e-5
    # an all-ones mask equals no mask (SURVEY.md 8a point 3)
    a = hubert_forward(sd, x[:1], None, 3).numpy()
    b = hubert_forward(sd, x[:1], [16000], 3).numpy()
    assert np.array_equal(a, b)
