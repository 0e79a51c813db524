"""Trimmed mode (`Segmenter(trim_padding=True)` / SYL_TRIM_PADDING): the padded frames of short
clips in a batch are not computed.  Documented deviation from sylber.py:93-126, where padded frames are computed,
returned and segmented: here hidden rows >= valid_frames[b] are zeros and carry no segments.  What must hold:
  * the VALID frames are bit-identical to the default (reference-padding) mode - same T_max, same GroupNorm statistics,
    same arithmetic per output element - and therefore within 1e-3 of the fp32 CPU oracle;
  * the on-device segmentation equals the oracle's segmentation of the returned (zero-tailed) states;
  * segments inside the valid region equal the default mode's up to the run that touches the valid / padding boundary."""
import numpy as np
import pytest
import torch

from oracle.hubert_ref import hubert_forward, num_frames
from oracle import segment_ref as R
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("mode", ["parity", "exact"])
def test_trimmed_mode_valid_frames_identical_padding_zero(cuda, mode):
    sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
    gen = torch.Generator().manual_seed(2)
    lens = [160000, 31234, 16000, 9000, 100000, 400, 160000]        # 400 samples = one frame
    wavs = [torch.randn(1, n, generator=gen) for n in lens]
    full = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", mode=mode)(wav=wavs, in_second=False)
    trim = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", mode=mode, trim_padding=True)
    outs = trim(wav=wavs, in_second=False)
    again = trim(wav=wavs, in_second=False)                          # CUDA-graph replay, stale workspace contents
    batch = torch.zeros(len(lens), max(lens))
    for i, w in enumerate(wavs):
        batch[i, :lens[i]] = w[0]
    ref = hubert_forward(sd, batch[:4], lens[:4], 9).numpy()
    t_max = num_frames(max(lens))
    for i, (o, f, a) in enumerate(zip(outs, full, again)):
        v = num_frames(lens[i])
        hs = o["hidden_states"]
        assert hs.shape == (t_max, 768)
        assert np.array_equal(hs[:v], f["hidden_states"][:v]), i           # valid frames: the same bits as the default mode
        assert not hs[v:].any(), i                                          # padded frames: zeros
        assert np.array_equal(hs, a["hidden_states"]) and np.array_equal(np.asarray(o["segments"]), np.asarray(a["segments"]))
        own = R.c_get_segment(hs, 2.6, 0.8)
        got = np.asarray(o["segments"]).reshape(-1, 2)
        assert len(own) == len(got) and (len(own) == 0 or np.array_equal(own, got))
        assert len(got) == 0 or got.max() <= v                              # no segment reaches into the padding
        if len(own):
            assert np.array_equal(R.c_segment_mean(hs, own), o["segment_features"], equal_nan=True)
        # inside the valid region the segments are the default mode's, except for the run that touches the boundary
        fs = np.asarray(f["segments"]).reshape(-1, 2)
        inner_full = [tuple(s) for s in fs if s[1] < v - 1]
        inner_trim = [tuple(s) for s in got if s[1] < v - 1]
        assert inner_full == inner_trim, i
        if i < 4:
            assert _rel(hs[:v], ref[i][:v]) < (1e-3 if mode != "exact" else 1e-4), (mode, i)


def test_trimmed_mode_without_padding_is_the_default_mode(cuda):
    sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
    wav = torch.randn(3, 48000, generator=torch.Generator().manual_seed(3))
    clips = [wav[i:i + 1] for i in range(3)]
    a = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0")(wav=clips, in_second=False)
    b = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", trim_padding=True)(wav=clips, in_second=False)
    for x, y in zip(a, b):
        assert np.array_equal(x["hidden_states"], y["hidden_states"])
        assert np.array_equal(np.asarray(x["segments"]), np.asarray(y["segments"]))
