"""GPU parity of the rows either side of the path: the PCM front door, the k-means lookup, the
`Sylber.segment` contract and length-bucketed batching, each against its CPU oracle."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_util as G
from oracle import frontdoor_ref, segment_ref as R
from seg_cases import plateau_states
from sylber_b200 import Segmenter, KMeansQuantizer, plan_length_buckets
from sylber_b200.weights import syllabic_test_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def seg9():
    return Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0), device="cuda:0")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _pcm(rng, n):
    t = np.arange(n) / 16000.0
    x = 4000 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + rng.normal(0, 800, n) + 150.0
    return np.clip(x, -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("lens", [[16000, 9000, 400], [160000], [33333, 160000, 8191, 8192, 8193]])
def test_prepare_pcm16_vs_oracle(lib, cuda, lens):
    rng = np.random.default_rng(len(lens))
    clips = [_pcm(rng, n) for n in lens]
    want, _ = frontdoor_ref.normalize_pcm16(clips)
    B, t_max = len(lens), max(lens)
    pcm = torch.from_numpy(np.concatenate(clips)).to(cuda)
    off = torch.tensor(np.concatenate([[0], np.cumsum(lens)[:-1]]), dtype=torch.int64, device=cuda)
    n = torch.tensor(lens, dtype=torch.int32, device=cuda)
    out = torch.full((B, t_max), 7.0, device=cuda)
    need = lib.syl_pcm16_workspace_bytes(B, t_max)
    ws = torch.empty(need, dtype=torch.uint8, device=cuda)
    for normalize in (1, 0):
        rc = lib.syl_prepare_pcm16(G.ptr(pcm), G.ptr(off), G.ptr(n), B, t_max, normalize, G.ptr(out), G.ptr(ws), need, G.stream())
        assert rc == 0, lib.syl_last_error(None)
        torch.cuda.synchronize()
        got = out.cpu()
        ref = want if normalize else frontdoor_ref.normalize_pcm16(clips, normalize=False)[0]
        # fp64 statistics on the device against torch's fp32 mean / std: agree to a few fp32 ulps of the result
        assert float((got - ref).abs().max()) < (2e-5 if normalize else 1e-7)
        for i, k in enumerate(lens):
            assert float(got[i, k:].abs().sum()) == 0.0


def test_segmenter_pcm16_branch_matches_wav_branch(seg9):
    rng = np.random.default_rng(7)
    clips = [_pcm(rng, n) for n in (32000, 20000, 9000)]
    want_wav, lens = frontdoor_ref.normalize_pcm16(clips)
    a = seg9(pcm16=clips, in_second=False)
    b = seg9(wav=[want_wav[i:i + 1, :lens[i]] for i in range(3)], in_second=False)
    assert isinstance(a, list) and len(a) == 3
    for x, y in zip(a, b):
        assert x["hidden_states"].shape == y["hidden_states"].shape
        # the two inputs differ by fp32 ulps (fp64 vs fp32 statistics); single-pass fp16 operands turn that into
        # different rounding decisions, so the outputs agree to the path's own error level, not to 1e-6
        assert _rel(x["hidden_states"], y["hidden_states"]) < 5e-4
    single = seg9(pcm16=clips[0])
    assert isinstance(single, dict) and single["hidden_states"].shape[1] == 768
    with pytest.raises(TypeError):
        seg9(pcm16=[clips[0].astype(np.float32)])


@pytest.mark.parametrize("n,K,normalize", [(300, 1000, False), (37, 4097, True), (1, 3, False), (8, 8, True)])
def test_kmeans_assign_vs_oracle(n, K, normalize):
    rng = np.random.default_rng(n + K)
    cent = (rng.normal(size=(K, 768)) * (6 / np.sqrt(768) if normalize else 1.0)).astype(np.float32)
    x = (cent[rng.integers(0, K, size=n)] + rng.normal(size=(n, 768)) * 0.3).astype(np.float32)
    q = KMeansQuantizer(cent, normalize=normalize)
    idx, dist = q.get_indices(x.reshape(n, 768), return_distance=True)
    want, d = frontdoor_ref.kmeans_assign(x, cent, normalize=normalize)
    assert idx.shape == (n, 1)                                  # trailing quantizer axis, as the reference's indices[0]
    idx = idx[:, 0].cpu().numpy()
    dmin = d.min(-1)
    chosen = d[np.arange(n), idx]
    assert np.all(chosen <= dmin * (1 + 1e-5) + 1e-4)          # identical except at numerical ties
    assert (idx == want).mean() >= 0.99
    assert np.allclose(dist.cpu().numpy(), chosen, rtol=1e-4, atol=1e-3)
    tok = q.get_indices(x.reshape(1, n, 768))
    assert tok.shape == (1, n, 1)
    assert torch.equal(q.decode(tok).cpu(), torch.from_numpy(cent[tok[..., 0].cpu().numpy()]))     # decode slices [..., :1]
    assert torch.equal(q.decode(torch.tensor([0, -1])).cpu(), torch.from_numpy(cent[[0, 0]]))
    assert q.get_indices(np.zeros((0, 768), np.float32)).shape == (0, 1)


def test_sylber_segment_contract_on_given_features(seg9):
    rng = np.random.default_rng(11)
    feats = np.stack([plateau_states(rng, 120) for _ in range(3)] + [np.zeros((120, 768), np.float32)])
    features, segments, avg = seg9.segment(features=torch.from_numpy(feats), mergethreshold=0.8, normthreshold=2.6)
    want_seg, want_avg = frontdoor_ref.sylber_segment(feats, 2.6, 0.8)
    assert features.is_cuda and tuple(features.shape) == feats.shape and torch.equal(features.cpu(), torch.from_numpy(feats))
    assert len(segments) == 4 and tuple(avg.shape) == want_avg.shape
    for got, want in zip(segments, want_seg):
        want = np.asarray(want)
        assert got.shape == want.shape and (want.size == 0 or np.array_equal(got, want.astype(np.int64)))
    assert segments[3].shape == (0,) and float(avg[3].abs().sum()) == 0.0       # empty utterance: one zero row
    assert np.allclose(avg.cpu().numpy(), want_avg, rtol=0, atol=2e-6)


def test_sylber_segment_contract_from_waveforms(seg9):
    g = torch.Generator().manual_seed(3)
    wav = torch.randn(2, 24000, generator=g)
    mask = torch.ones(2, 24000, dtype=torch.long)
    mask[1, 16000:] = 0
    wav[1, 16000:] = 0
    features, segments, avg = seg9.segment(input_values=wav, attention_mask=mask)
    ref = seg9.speech_model(wav, attention_mask=mask).last_hidden_state
    assert torch.equal(features, ref)
    f = features.cpu().numpy()
    for b in range(2):
        want = R.c_get_segment(f[b], 2.6, 0.8)
        assert np.array_equal(np.asarray(segments[b]).reshape(-1, 2), np.asarray(want).reshape(-1, 2))
        if len(want):
            assert np.array_equal(avg[b, :len(want)].cpu().numpy(), R.c_segment_mean(f[b], want))


def test_bucketed_batching_equals_per_bucket_calls(seg9):
    g = torch.Generator().manual_seed(5)
    lens = [48000, 16000, 46000, 17000, 30000]
    wavs = [torch.randn(1, n, generator=g) for n in lens]
    sd = syllabic_test_state_dict(9, 0)
    bucketed = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", bucket_ratio=1.25)
    got = bucketed(wav=wavs, in_second=False)
    plan = plan_length_buckets(lens, 1.25, 64)
    assert sorted(map(tuple, plan)) == [(0, 2), (1, 3), (4,)]
    for idx in plan:
        want = seg9(wav=[wavs[i] for i in idx], in_second=False)
        for i, w in zip(idx, want):
            assert got[i]["hidden_states"].shape == w["hidden_states"].shape        # padded to the BUCKET's maximum
            assert np.array_equal(got[i]["hidden_states"], w["hidden_states"])
            assert np.array_equal(np.asarray(got[i]["segments"]), np.asarray(w["segments"]))


@pytest.mark.parametrize("rate,lens", [(44100, [44100, 30000]), (8000, [8000, 12001, 400]), (48000, [96000])])
def test_resample_vs_torchaudio(lib, cuda, rate, lens):
    import math
    torchaudio = pytest.importorskip("torchaudio")
    from sylber_b200.resample import sinc_resample_kernel, resampled_length
    k, width, orig_g, new_g = sinc_resample_kernel(rate, 16000)
    g = torch.Generator().manual_seed(rate)
    B, t_in = len(lens), max(lens)
    x = torch.zeros(B, t_in)
    for i, n in enumerate(lens):
        x[i, :n] = torch.randn(n, generator=g)
    n_out = [resampled_length(n, orig_g, new_g) for n in lens]
    t_out = max(n_out)
    out = torch.full((B, t_out), 7.0, device=cuda)
    n_dev = torch.zeros(B, dtype=torch.int32, device=cuda)
    x_dev, n_in, k_dev = x.to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), torch.from_numpy(k).to(cuda)   # keep alive
    rc = lib.syl_resample(G.ptr(x_dev), G.ptr(n_in), B, t_in, G.ptr(k_dev), orig_g, new_g, width, G.ptr(out), G.ptr(n_dev), t_out,
                          G.stream())
    assert rc == 0, lib.syl_last_error(None)
    torch.cuda.synchronize()
    assert n_dev.cpu().tolist() == n_out, (n_dev.cpu().tolist(), n_out)
    got = out.cpu()
    for i, n in enumerate(lens):
        # oracle: the same filter bank evaluated in float64 (pinned against torchaudio on the CPU side); the kernel's
        # fp32 FMA chain over up to 475 taps stays within a few 1e-6 of it.  torchaudio's own fp32 result is a second
        # fp32 implementation (~1e-5 from the exact sum), so it only gets a loose bound here.
        exact = frontdoor_ref.resample_f64(x[i, :n].numpy(), rate)
        assert exact.shape[0] == n_out[i]
        diff = np.abs(got[i, :n_out[i]].numpy().astype(np.float64) - exact)
        j = int(diff.argmax())
        # (the message carries the evidence: this test failed in 2 of ~10 full-suite runs in round 2, never in isolation,
        # with the kernel's arithmetic 4e-7 from the exact sum in emulation - see profiles/r02_next_steps.md)
        assert diff[j] < 2e-5, f"clip {i}: |gpu - exact| = {diff[j]:.3e} at sample {j} of {n_out[i]} (gpu {got[i, j]:.6f}, exact {exact[j]:.6f}); " \
                               f"{int((diff > 2e-5).sum())} samples off"
        # torchaudio's conv1d is whatever fp32 kernel oneDNN picks for the host CPU, and the boxes differ: only gross errors
        # (a wrong phase or tap) are checked against it here; tests/test_oracle_frontdoor.py pins the oracle to it
        want = torchaudio.functional.resample(x[i:i + 1, :n], rate, 16000)[0]
        assert float((got[i, :n_out[i]] - want).abs().max()) < 1e-3
        assert float(got[i, n_out[i]:].abs().sum()) == 0.0, f"clip {i}: tail not zero"


def test_segmenter_pcm16_other_sample_rate(seg9):
    rng = np.random.default_rng(9)
    clips = [_pcm(rng, n) for n in (44100 * 2, 30000)]
    want_wav, lens = frontdoor_ref.normalize_pcm16(clips, sample_rate=44100)
    a = seg9(pcm16=clips, sample_rate=44100, in_second=False)
    b = seg9(wav=[want_wav[i:i + 1, :lens[i]] for i in range(2)], in_second=False)
    for x, y in zip(a, b):
        assert x["hidden_states"].shape == y["hidden_states"].shape
        assert _rel(x["hidden_states"], y["hidden_states"]) < 5e-4


@pytest.mark.parametrize("rate", [16000, 22050])
def test_wav_file_branch_runs_its_preprocessing_on_the_device(seg9, tmp_path, rate):
    """`Segmenter(wav_file)` - the reference's file branch (sylber.py:83-87): read, resample to 16 kHz, (w - mean) / std.
    Mono files of one sample rate keep only the file read on the host; the rest is the device front door, so the result
    is bit-identical to the `pcm16=` call on the same samples, and within the path's own error of the host-normalised
    `wav=` call."""
    from scipy.io import wavfile
    rng = np.random.default_rng(11)
    clips = [_pcm(rng, n) for n in (rate * 2, rate + 1234)]
    paths = []
    for i, c in enumerate(clips):
        path = tmp_path / f"clip{i}.wav"
        wavfile.write(str(path), rate, c)
        paths.append(str(path))
    a = seg9(wav_file=paths, in_second=False)
    b = seg9(pcm16=clips, sample_rate=rate, in_second=False)
    want_wav, lens = frontdoor_ref.normalize_pcm16(clips, sample_rate=rate)
    c = seg9(wav=[want_wav[i:i + 1, :lens[i]] for i in range(2)], in_second=False)
    assert isinstance(a, list) and len(a) == 2
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x["hidden_states"], y["hidden_states"])
        assert np.array_equal(np.asarray(x["segments"]), np.asarray(y["segments"]))
        assert _rel(x["hidden_states"], z["hidden_states"]) < 5e-4
    single = seg9(paths[0])                      # first positional argument is a path, dict out (sylber.py:63,138)
    assert isinstance(single, dict) and np.array_equal(single["hidden_states"][:10], seg9(pcm16=clips[0], sample_rate=rate)["hidden_states"][:10])
