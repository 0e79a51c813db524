"""CPU: the host logic and oracles of the rows either side of the path: length bucketing, the
Thresholder port against the reference class, the PCM normalisation oracle against the reference's own lines."""
import os
import sys

import numpy as np
import pytest
import torch

from sylber_b200.batching import plan_length_buckets, padded_work
from sylber_b200.thresholder import Thresholder
from oracle import frontdoor_ref

REF = "/root/reference"


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bucket_plan_is_a_partition_within_ratio(seed):
    rng = np.random.default_rng(seed)
    lengths = rng.integers(32000, 480001, size=97).tolist()
    for ratio, max_batch in ((1.25, 64), (1.0, 8), (2.0, 5)):
        buckets = plan_length_buckets(lengths, ratio, max_batch)
        flat = sorted(i for b in buckets for i in b)
        assert flat == list(range(len(lengths)))
        for b in buckets:
            assert 1 <= len(b) <= max_batch and b == sorted(b)
            ls = [lengths[i] for i in b]
            assert max(ls) <= min(ls) * ratio + 1e-9
        assert plan_length_buckets(lengths, ratio, max_batch) == buckets       # deterministic
        assert padded_work(lengths, buckets) <= padded_work(lengths)
    # config 5's distribution (2-30 s): bucketing removes most of the padding
    assert padded_work(lengths, plan_length_buckets(lengths, 1.25, 64)) < 0.75 * padded_work(lengths)


def test_bucket_plan_edge_cases():
    assert plan_length_buckets([], 1.25, 4) == []
    assert plan_length_buckets([5], 1.25, 4) == [[0]]
    assert plan_length_buckets([10, 10, 10], 1.0, 2) == [[0, 1], [2]]
    with pytest.raises(ValueError):
        plan_length_buckets([1, 2], 0.5, 4)


# known answers computed with the unmodified reference class (torch 2.11, float32), see the test below
THRESHOLD_CASES = [
    ((8.0, 4.0, 1.5, 0.25), 3.0015573501586914),
    ((6.5, 9.0, 2.0, 1.0), 3.7439115047454834),
    ((3.0, 1.0, 1.0, 1.0), 2.0),
    ((3.0, 2.0, 1.0, 1.0), 2.063706159591675),
]


def test_thresholder_known_answers():
    for (sm, sv, nm, nv), want in THRESHOLD_CASES:
        got = Thresholder(sm, sv, nm, nv).get_threshold()
        assert abs(float(got) - want) <= 2e-6 * abs(want), (sm, sv, nm, nv, got, want)
    assert float(Thresholder(threshold=2.6).get_threshold()) == pytest.approx(2.6, rel=1e-7)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this box")
def test_thresholder_matches_reference_class():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_segment_utils", os.path.join(REF, "sylber/utils/segment_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    RefThresholder = mod.Thresholder
    rng = np.random.default_rng(0)
    for (sm, sv, nm, nv), want in THRESHOLD_CASES:
        ref = RefThresholder(sm, sv, nm, nv, decay=0.99)
        ours = Thresholder(sm, sv, nm, nv, decay=0.99)
        assert float(ref.get_threshold()) == pytest.approx(want, rel=1e-6)
        for _ in range(5):
            sig = rng.normal(sm, 2.0, size=300).astype(np.float32)
            noi = rng.normal(nm, 0.5, size=200).astype(np.float32)
            ref.update_stats(torch.from_numpy(sig), torch.from_numpy(noi))
            ours.update_stats(sig, noi)
            assert float(ours.get_threshold()) == pytest.approx(float(ref.get_threshold()), rel=2e-6)
            assert float(ours.signal_var) == pytest.approx(float(ref.signal_var), rel=1e-6)


def test_normalize_pcm16_oracle_follows_reference_lines():
    rng = np.random.default_rng(3)
    pcm = [(rng.normal(0, 3000, size=n)).astype(np.int16) for n in (16000, 9000)]
    out, lens = frontdoor_ref.normalize_pcm16(pcm)
    assert lens == [16000, 9000] and out.shape == (2, 16000)
    for i, x in enumerate(pcm):
        w = torch.from_numpy(x.astype(np.float32) / 32768.0)[None]          # torchaudio.load of 16-bit PCM
        w = (w - w.mean()) / w.std()                                        # sylber.py:86
        assert torch.equal(out[i, :lens[i]], w[0])
        assert float(out[i, lens[i]:].abs().sum()) == 0.0
        assert abs(float(out[i, :lens[i]].mean())) < 1e-5 and float(out[i, :lens[i]].std()) == pytest.approx(1.0, abs=1e-5)


def test_kmeans_oracle_matches_brute_force():
    rng = np.random.default_rng(4)
    c = rng.normal(size=(50, 768)).astype(np.float32)
    x = rng.normal(size=(20, 768)).astype(np.float32)
    idx, _ = frontdoor_ref.kmeans_assign(x, c)
    brute = np.array([np.argmin(((xi[None].astype(np.float64) - c.astype(np.float64)) ** 2).sum(-1)) for xi in x])
    assert np.array_equal(idx, brute)
    idx_n, _ = frontdoor_ref.kmeans_assign(x * 3.0, c, normalize=True)
    idx_n2, _ = frontdoor_ref.kmeans_assign(x * 7.0, c, normalize=True)
    assert np.array_equal(idx_n, idx_n2)          # the normalisation removes the scale


@pytest.mark.parametrize("rate", [44100, 48000, 8000, 22050, 11025])
def test_resample_filter_bank_equals_torchaudio(rate):
    """sylber_b200/resample.py restates torchaudio's sinc_interp_hann kernel (the reference calls
    torchaudio.transforms.Resample(sr, 16000), sylber.py:85); it must be the same filter, bit for bit."""
    import math
    torchaudio = pytest.importorskip("torchaudio")
    from sylber_b200.resample import sinc_resample_kernel, resampled_length
    g = math.gcd(rate, 16000)
    want, width = torchaudio.functional.functional._get_sinc_resample_kernel(rate, 16000, g)
    k, w, orig_g, new_g = sinc_resample_kernel(rate, 16000)
    assert (w, orig_g, new_g) == (width, rate // g, 16000 // g)
    assert np.array_equal(k, want[:, 0, :].numpy())
    x = torch.randn(1, 12345)
    assert torchaudio.functional.resample(x, rate, 16000).shape[1] == resampled_length(12345, orig_g, new_g)


def test_sub_batch_bounds_cover_every_row_once():
    from sylber_b200.batching import sub_batch_bounds
    for n in list(range(1, 70)) + [100, 256]:
        for streams in (1, 2, 3, 4):
            for max_batch in (32, 64):
                b = sub_batch_bounds(n, streams, max_batch)
                assert b[0][0] == 0 and b[-1][1] == n
                assert all(lo < hi for lo, hi in b) and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
                assert all(hi - lo <= max_batch for lo, hi in b)
                if n <= max_batch:
                    assert len(b) <= streams
    assert [hi - lo for lo, hi in sub_batch_bounds(32, 3, 32)] == [8, 12, 12]
    assert [hi - lo for lo, hi in sub_batch_bounds(7, 3, 32)] == [7]
    assert [hi - lo for lo, hi in sub_batch_bounds(80, 3, 32)] == [32, 32, 16]
    assert sub_batch_bounds(32, 3, 32, explicit=[16, 16]) == [(0, 16), (16, 32)]
    assert [hi - lo for lo, hi in sub_batch_bounds(32, 3, 32, explicit=[5, 5])] == [8, 12, 12]      # wrong sum: rule applies


def test_resampled_length_matches_ceil():
    import math
    from sylber_b200.resample import resampled_length
    for rate in (8000, 11025, 22050, 44100, 48000):
        g = math.gcd(rate, 16000)
        for n in (1, 2, 399, 400, 12345, 441000):
            assert resampled_length(n, rate // g, 16000 // g) == math.ceil(16000 * n / rate)


@pytest.mark.parametrize("rate,n", [(44100, 44100), (8000, 12001), (48000, 30000), (22050, 400)])
def test_resample_f64_oracle_agrees_with_torchaudio(rate, n):
    torchaudio = pytest.importorskip("torchaudio")
    x = torch.randn(n, generator=torch.Generator().manual_seed(rate + n))
    want = torchaudio.functional.resample(x[None], rate, 16000)[0].numpy()
    got = frontdoor_ref.resample_f64(x.numpy(), rate)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 3e-5          # torchaudio's fp32 conv1d against the exact sum
