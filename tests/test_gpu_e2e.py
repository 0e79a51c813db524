"""End-to-end parity of the drop-in Segmenter on a real B200 against the CPU oracle and the reference's golden
outputs.  Tolerance (BASELINE.json north_star): segments identical, hidden_states / segment_features within 1e-3
relative (Frobenius) of the fp32 reference."""
import os

import numpy as np
import pytest
import torch

import gpu_util as G
from oracle.hubert_ref import hubert_forward, num_frames
from oracle import segment_ref as R
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, random_hubert_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def seg9():
    return Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0), device="cuda:0")


def _check_against_own_states(out):
    """The on-device segmentation / pooling must equal the oracle run on the SAME hidden states, bit for bit."""
    hs = out["hidden_states"]
    want = R.c_get_segment(hs, 2.6, 0.8)
    got = np.asarray(out["segments"]).reshape(-1, 2)
    assert len(want) == len(got) and (len(want) == 0 or np.array_equal(want, got.astype(np.int64)))
    if len(want):
        assert np.array_equal(R.c_segment_mean(hs, want), out["segment_features"], equal_nan=True)


def test_config1_sample_wav_vs_reference_golden(seg9):
    g = np.load(os.path.join(GOLD, "sample_wav.npz"))
    out = seg9(wav=torch.from_numpy(g["wav"]), in_second=False)
    assert isinstance(out, dict) and set(out) == {"segments", "segment_features", "hidden_states"}
    assert out["hidden_states"].shape == (143, 768) and out["hidden_states"].dtype == np.float32
    assert _rel(out["hidden_states"], g["hidden_states"]) < TOL
    _check_against_own_states(out)
    assert out["segments"].dtype == np.int64
    assert np.array_equal(out["segments"], g["segments"])            # identical frame indices vs the reference
    assert _rel(out["segment_features"], g["segment_features"]) < TOL
    sec = seg9(wav=torch.from_numpy(g["wav"]), in_second=True)
    assert sec["segments"].dtype == np.float64 and np.array_equal(sec["segments"], g["segments_sec"])


def test_padded_list_input_vs_reference_golden(seg9):
    g = np.load(os.path.join(GOLD, "sample_wav.npz"))
    outs = seg9(wav=[torch.from_numpy(g["b_wav0"]), torch.from_numpy(g["b_wav1"])], in_second=False)
    assert isinstance(outs, list) and len(outs) == 2
    for i, o in enumerate(outs):
        assert o["hidden_states"].shape == g[f"b_hidden{i}"].shape        # T_max rows, padded frames included
        assert _rel(o["hidden_states"], g[f"b_hidden{i}"]) < TOL
        _check_against_own_states(o)
        assert np.array_equal(o["segments"], g[f"b_segments{i}"])
        assert _rel(o["segment_features"], g[f"b_features{i}"]) < TOL


def test_mixed_lengths_vs_oracle_and_padding_semantics(seg9):
    sd = syllabic_test_state_dict(9, 0)
    gen = torch.Generator().manual_seed(2)
    lens = [48000, 31234, 16000, 9000, 48000]
    wavs = [torch.randn(1, n, generator=gen) for n in lens]
    outs = seg9(wav=wavs, in_second=False)
    batch = torch.zeros(len(lens), max(lens))
    for i, w in enumerate(wavs):
        batch[i, :w.shape[1]] = w[0]
    ref = hubert_forward(sd, batch, lens, 9).numpy()
    for i, o in enumerate(outs):
        assert _rel(o["hidden_states"], ref[i]) < TOL
        _check_against_own_states(o)
    # an utterance's result depends on its own samples and T_max only: alone-but-padded == in-batch
    alone = torch.zeros(1, max(lens))
    alone[0, :lens[3]] = wavs[3][0]
    eng = seg9._engine
    n = torch.tensor([lens[3]], dtype=torch.int32, device=eng.device)
    hid, _, _, _ = eng.forward(alone.to(eng.device), n, 2.6, 0.8, segment=False)
    assert np.array_equal(hid[0].cpu().numpy(), outs[3]["hidden_states"])
    # determinism: same input twice -> identical bits
    again = seg9(wav=wavs, in_second=False)
    for a, b in zip(outs, again):
        assert np.array_equal(a["hidden_states"], b["hidden_states"]) and np.array_equal(a["segments"], b["segments"])


def test_empty_result_contract(seg9):
    """Plain random weights give frame norms ~27.7, so a norm threshold above that yields no segments."""
    s = Segmenter(model_ckpt=None, state_dict=random_hubert_state_dict(2, 1), encoding_layer=2, norm_threshold=1e4,
                  device="cuda:0")
    out = s(wav=torch.randn(1, 8000, generator=torch.Generator().manual_seed(0)))
    assert out["segments"].shape == (0,) and out["segments"].dtype == np.float64      # np.array([]) * 1.0 / 50
    assert out["segment_features"].shape == (0,)
    assert out["hidden_states"].shape == (num_frames(8000), 768)


def test_modes_and_per_stage_drift():
    sd = syllabic_test_state_dict(9, 0)
    gen = torch.Generator().manual_seed(1)
    lens = [48000, 30000]
    batch = torch.zeros(2, 48000)
    for i, n in enumerate(lens):
        batch[i, :n] = torch.randn(n, generator=gen)
    stages = {}
    ref = hubert_forward(sd, batch, lens, 9, stages=stages).numpy()
    errs = {}
    for mode in ("parity", "strict", "fast", "exact"):
        s = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", mode=mode)
        eng = s._engine
        hid, _, _, _ = eng.forward(batch.to(eng.device), torch.tensor(lens, dtype=torch.int32, device=eng.device), 2.6, 0.8,
                                   segment=False)
        errs[mode] = _rel(hid.cpu().numpy(), ref)
        if mode == "strict":
            for i in range(7):
                L = stages[f"conv{i}"].shape[2]
                got = eng.read_stage(f"conv{i}", (2, L, 512)).cpu().numpy()
                assert _rel(got, stages[f"conv{i}"].transpose(1, 2).numpy()) < 2e-4, i
            got = eng.read_stage("pos", (2, hid.shape[1], 768)).cpu().numpy()
            assert _rel(got, stages["pos"].numpy()) < 2e-4
    assert errs["parity"] < TOL / 2 and errs["strict"] < TOL / 2 and errs["exact"] < 1e-4
    assert errs["exact"] < errs["strict"] <= errs["parity"] * 1.05 and errs["parity"] <= errs["fast"] * 1.05


def test_speech_model_seam(seg9):
    """`.speech_model(input_values, attention_mask=...)` returns an object with `.last_hidden_state` (sylber.py:122)."""
    sd = syllabic_test_state_dict(9, 0)
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(2, 20000, generator=gen)
    x[1, 12000:] = 0
    mask = torch.ones(2, 20000, dtype=torch.long)
    mask[1, 12000:] = 0
    out = seg9.speech_model(x, attention_mask=mask).last_hidden_state
    assert out.is_cuda and out.shape == (2, num_frames(20000), 768)
    ref = hubert_forward(sd, x, [20000, 12000], 9).numpy()
    assert _rel(out.cpu().numpy(), ref) < TOL
    assert seg9.enc_dim == 768 and seg9.encoding_layer == 9 and seg9.norm_threshold == 2.6 and seg9.merge_threshold == 0.8


def test_full_size_config2_properties(seg9):
    """BASELINE config 2 (batch 32 x 10 s): size-independent properties instead of a full CPU oracle run."""
    gen = torch.Generator().manual_seed(1)
    wav = torch.randn(32, 160000, generator=gen)
    eng = seg9._engine
    n = torch.full((32,), 160000, dtype=torch.int32, device=eng.device)
    hid, seg, cnt, feat = eng.forward(wav.to(eng.device), n, np.float32(2.6), np.float32(0.8))
    torch.cuda.synchronize()
    assert hid.shape == (32, 499, 768) and bool(torch.isfinite(hid).all())
    hs = hid.cpu().numpy()
    cnt_h, seg_h, feat_h = cnt.cpu().numpy(), seg.cpu().numpy(), feat.cpu().numpy()
    for b in (0, 13, 31):                                    # on-device segmentation == oracle on the same states
        want = R.c_get_segment(hs[b], 2.6, 0.8)
        assert np.array_equal(want, seg_h[b, :cnt_h[b]].astype(np.int64))
        assert np.array_equal(R.c_segment_mean(hs[b], want), feat_h[b, :cnt_h[b]], equal_nan=True)
    s = seg_h[0, :cnt_h[0]]
    assert (s[:, 0] < s[:, 1]).all() and (s[1:, 0] >= s[:-1, 1]).all() and s.max() <= 499   # sorted, disjoint, in range
    # batch invariance: rows 0..3 alone give the same bits as inside the batch of 32
    hid4, _, _, _ = eng.forward(wav[:4].to(eng.device), n[:4], 2.6, 0.8, segment=False)
    assert np.array_equal(hid4.cpu().numpy(), hs[:4])
    # one full-size row against the CPU oracle
    ref = hubert_forward(syllabic_test_state_dict(9, 0), wav[:1], [160000], 9).numpy()
    assert _rel(hs[0], ref[0]) < TOL


def test_long_form_60s(seg9):
    """BASELINE config 4 shape (60 s, T = 2999) on one clip: attention over 24 key blocks vs the CPU oracle."""
    gen = torch.Generator().manual_seed(3)
    wav = torch.randn(1, 960000, generator=gen)
    out = seg9(wav=wav, in_second=False)
    assert out["hidden_states"].shape == (2999, 768)
    ref = hubert_forward(syllabic_test_state_dict(9, 0), wav, [960000], 9).numpy()
    assert _rel(out["hidden_states"], ref[0]) < TOL
    _check_against_own_states(out)


def test_twelve_layer_encoder_vs_oracle():
    """BASELINE.json names the 12-layer hubert-base encoder (the reference checkpoint keeps 9): same path, n_layers = 12,
    on a padded 3-clip batch against the CPU oracle, in every precision preset."""
    sd = syllabic_test_state_dict(12, 1)
    gen = torch.Generator().manual_seed(12)
    lens = [40000, 27000, 16000]
    wavs = [torch.randn(1, n, generator=gen) for n in lens]
    batch = torch.zeros(3, max(lens))
    for i, w in enumerate(wavs):
        batch[i, :lens[i]] = w[0]
    ref = hubert_forward(sd, batch, lens, 12).numpy()
    for mode, bar in (("parity", TOL), ("fast", TOL), ("exact", 1e-4)):
        seg = Segmenter(model_ckpt=None, state_dict=sd, encoding_layer=12, device="cuda:0", mode=mode)
        outs = seg(wav=wavs, in_second=False)
        assert seg.speech_model.config.num_hidden_layers == 12
        for i, o in enumerate(outs):
            assert o["hidden_states"].shape == ref[i].shape
            assert _rel(o["hidden_states"], ref[i]) < bar, (mode, i)
            _check_against_own_states(o)
        del seg


def test_hidden_states_can_stay_on_the_device(seg9):
    """`hidden_to="device"` / None (extension): the hidden states are not copied to the host; everything else is unchanged."""
    gen = torch.Generator().manual_seed(8)
    wavs = [torch.randn(1, n, generator=gen) for n in (40000, 16000, 30000)]
    host = seg9(wav=wavs, in_second=False)
    dev = seg9(wav=wavs, in_second=False, hidden_to="device")
    none = seg9(wav=wavs, in_second=False, hidden_to=None)
    for h, d, n in zip(host, dev, none):
        assert torch.is_tensor(d["hidden_states"]) and d["hidden_states"].is_cuda
        assert np.array_equal(d["hidden_states"].cpu().numpy(), h["hidden_states"])
        assert n["hidden_states"] is None
        for o in (d, n):
            assert np.array_equal(np.asarray(o["segments"]), np.asarray(h["segments"]))
            assert np.array_equal(np.asarray(o["segment_features"]), np.asarray(h["segment_features"]))
