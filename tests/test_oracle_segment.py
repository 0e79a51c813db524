"""The segmentation oracles (NumPy restatement + C twin) against the reference's golden vectors and, when the
reference checkout is present (build container only), against the live reference function."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import segment_ref as R
from seg_cases import plateau_states

GOLD = os.path.join(os.path.dirname(__file__), "golden", "segment_cases.npz")
REF_FILE = "/root/reference/sylber/utils/segment_utils.py"


def _same_segments(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0 or b.size == 0:
        return a.size == b.size
    return a.shape == b.shape and np.array_equal(a, b)


def test_golden_cases_numpy_and_c():
    g = np.load(GOLD)
    n = len([k for k in g.files if k.startswith("states_")])
    assert n >= 8
    for i in range(n):
        st, want, feat = g[f"states_{i}"], g[f"segments_{i}"], g[f"features_{i}"]
        got_np = R.get_segment(st, 2.6, 0.8)
        got_c = R.c_get_segment(st, 2.6, 0.8)
        assert _same_segments(got_np, want), i
        assert _same_segments(got_c, want), i
        if want.size:
            assert got_np.dtype == np.int64 and got_np.shape[1] == 2
            assert np.array_equal(R.c_segment_mean(st, want), feat, equal_nan=True)
            assert np.array_equal(R.package(st, got_np, False)["segment_features"], feat, equal_nan=True)
        else:
            # the reference returns np.array([]) - shape (0,), float64 - for an utterance without segments
            assert got_np.shape == (0,) and got_np.dtype == np.float64
            assert R.package(st, got_np)["segment_features"].shape == (0,)


def test_package_seconds():
    g = np.load(GOLD)
    st, seg = g["states_4"], g["segments_4"]
    out = R.package(st, seg, in_second=True)
    assert out["segments"].dtype == np.float64
    assert np.array_equal(out["segments"], seg * 1.0 / 50)
    assert out["hidden_states"] is st


@pytest.mark.parametrize("seed", [0, 1])
def test_numpy_vs_c_randomized(seed):
    rng = np.random.default_rng(seed)
    kinds = set()
    for _ in range(60):
        T = int(rng.integers(1, 260))
        st = plateau_states(rng, T)
        a = R.get_segment(st, 2.6, 0.8)
        b = R.c_get_segment(st, 2.6, 0.8)
        assert _same_segments(a, b)
        kinds.add(len(a) > 0)
        if len(a):
            f = np.stack([st[s:e].mean(0) for s, e in a])
            assert np.array_equal(f, R.c_segment_mean(st, a), equal_nan=True)
    assert kinds == {True} or kinds == {True, False}


def test_edge_cases():
    rng = np.random.default_rng(3)
    d = 768
    quiet = rng.standard_normal((20, d)).astype(np.float32) * 0.01          # all frames under the norm threshold
    assert R.c_get_segment(quiet, 2.6, 0.8).shape == (0, 2)
    assert R.get_segment(quiet, 2.6, 0.8).shape == (0,)
    c = rng.standard_normal(d).astype(np.float32)
    c *= 3.0 / np.linalg.norm(c)
    flat = np.repeat(c[None], 30, 0)                                          # one long segment
    assert R.c_get_segment(flat, 2.6, 0.8).tolist() == [[0, 30]]
    assert R.get_segment(flat, 2.6, 0.8).tolist() == [[0, 30]]
    one = flat[:1]
    assert R.c_get_segment(one, 2.6, 0.8).tolist() == [[0, 1]]
    ortho = np.zeros((6, d), np.float32)                                      # every frame its own segment
    for i in range(6):
        ortho[i, i] = 3.0
    a, b = R.get_segment(ortho, 2.6, 0.8), R.c_get_segment(ortho, 2.6, 0.8)
    assert _same_segments(a, b) and len(a) >= 5
    # thresholds compare in float32 (NumPy 2 weak scalars): a norm of exactly float32(2.6) is "on"
    v = np.zeros((1, d), np.float32)
    v[0, 0] = np.float32(2.6)
    assert len(R.get_segment(v, 2.6, 0.8)) == len(R.c_get_segment(v, 2.6, 0.8))


def test_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(11)
    for n in [0, 1, 7, 8, 9, 15, 16, 127, 128, 129, 255, 768, 769, 1500]:
        for _ in range(20):
            a = (rng.standard_normal(n) * rng.choice([1e-3, 1.0, 1e3])).astype(np.float32)
            assert R.c_np_sum(a) == a.sum(), n


def test_powf_is_not_sqrt_and_oracle_uses_powf():
    """np.float32 ** .5 is libm powf, which is not correctly rounded; the C oracle must call the same function."""
    x = np.float32(float.fromhex("0x1.66bf82p+8"))      # found by exhaustive search: powf(x,.5f) != sqrtf(x)
    lib = R._lib()
    assert np.float32(lib.syl_oracle_powf_half(float(x))) == x ** .5
    assert x ** .5 != np.sqrt(x)


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference checkout not present (GPU box)")
def test_against_live_reference():
    spec = importlib.util.spec_from_file_location("ref_segment_utils", REF_FILE)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(5)
    for _ in range(40):
        st = plateau_states(rng, int(rng.integers(1, 200)))
        want = ref.get_segment(st, 2.6, 0.8)
        assert _same_segments(R.get_segment(st, 2.6, 0.8), want)
        assert _same_segments(R.c_get_segment(st, 2.6, 0.8), want)


# ------------------------------------------------------------------------------------------------
# The CUDA scan is parallel over RUNS (csrc/segment.cuh): CTA c owns the runs that start in frames [16 c, 16 c + 16),
# keeps its k-th segment in slot (first run start + k) of a frame-indexed table, and the table is packed in frame
# order.  This CPU model of that decomposition (the oracle applied to each CTA's frame range) must equal the oracle
# on the whole utterance: it is the property the kernel's structure rests on (segment_utils.py:83-89: a masked-off
# frame resets the scan; :110-128: a mid-boundary only touches the two segments of one run).
# ------------------------------------------------------------------------------------------------
def _cta_ranges(on, chunk):
    T = len(on)
    for c0 in range(0, T, chunk):
        starts = [i for i in range(c0, min(T, c0 + chunk)) if on[i] and (i == 0 or not on[i - 1])]
        if not starts:
            continue
        end = starts[-1] + 1
        while end < T and on[end]:
            end += 1
        yield starts[0], end


def _run_parallel_model(states, norm_thr, merge_thr, chunk):
    T = len(states)
    on = ((states ** 2).sum(-1) + 1e-8) ** .5 >= norm_thr
    slot_s = np.zeros(T, np.int64)
    slot_e = np.full(T, -1, np.int64)             # < 0: unused (the kernel's SEG_SLOT_UNUSED / absorbed marker)
    for f0, f1 in _cta_ranges(on, chunk):
        runs, splits = R._scan(states[f0:f1], on[f0:f1], merge_thr)
        assert f0 + len(runs) <= f1               # phase-1 slots stay inside the CTA's own frames
        for k, (s, e) in enumerate(R._refine(states[f0:f1], runs, splits, merge_thr)):
            assert slot_e[f0 + k] < 0
            slot_s[f0 + k], slot_e[f0 + k] = s + f0, e + f0
    keep = slot_e >= 0
    return np.stack([slot_s[keep], slot_e[keep]], 1)


@pytest.mark.parametrize("chunk", [4, 16, 32])
def test_runs_are_independent(chunk):
    rng = np.random.default_rng(5)
    for _ in range(60):
        T = int(rng.integers(1, 300))
        st = plateau_states(rng, T, noise=float(rng.uniform(0.1, 0.6)), sil=float(rng.uniform(0.0, 0.5)))
        want = np.asarray(R.get_segment(st, 2.6, 0.8)).reshape(-1, 2)
        got = _run_parallel_model(st, 2.6, 0.8, chunk)
        assert want.shape == got.shape and np.array_equal(want, got)
