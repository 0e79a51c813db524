"""Per-kernel parity on a real B200, through the C ABI (include/sylber_b200.h)."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_util as G
from oracle import segment_ref as R
from seg_cases import long_segment_states, plateau_states

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 128), (1000, 768, 768), (4096, 2304, 768), (2000, 768, 3072)])
def test_gemm_vs_fp64(lib, cuda, M, N, K):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=cuda)
    W = torch.randn(N, K, device=cuda) * 0.05
    bias = torch.randn(N, device=cuda)
    ref = A.double() @ W.double().t() + bias.double()
    one = G.gemm_f32(lib, A, W, bias=bias, n_pass=1)
    three = G.gemm_f32(lib, A, W, bias=bias, n_pass=3)
    assert G.rel_err(one, ref) < 6e-4          # fp16 operands, fp32 accumulate: ~2.5e-4 measured
    assert G.rel_err(three, ref) < 4e-5        # hi/lo split operands: fp32-class, ~3e-6..1.5e-5 measured
    assert G.rel_err(three, ref) < G.rel_err(one, ref) / 8


def test_gemm_gelu_epilogue(lib, cuda):
    torch.manual_seed(0)
    A = torch.randn(777, 512, device=cuda)
    W = torch.randn(512, 512, device=cuda) * 0.06
    bias = torch.randn(512, device=cuda) * 0.3
    res = torch.randn(777, 512, device=cuda)
    out = G.gemm_f32(lib, A, W, bias=bias, residual=res, n_pass=3, act=1)
    ref = torch.nn.functional.gelu(A.double() @ W.double().t() + bias.double()) + res.double()
    assert G.rel_err(out, ref) < 2e-5


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,T,lens", [(2, 499, None), (3, 143, [143, 100, 17]), (1, 1000, [777]), (2, 130, [1, 129]),
                                      (1, 2999, None)])
def test_attention_vs_torch(lib, cuda, B, T, lens):
    torch.manual_seed(T)
    qkv = torch.randn(B * T, 2304, device=cuda)
    qkv[:, :768] *= 0.125 * 1.5                 # Q arrives pre-scaled by 1/sqrt(64)
    q16 = qkv.half()
    out = torch.zeros(B * T, 768, dtype=torch.float16, device=cuda)
    kv = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=cuda)
    rc = lib.syl_attention(G.ptr(q16), G.ptr(kv), B, T, G.ptr(out), G.stream())
    assert rc == 0, lib.syl_last_error(None)
    torch.cuda.synchronize()
    x = q16.float().view(B, T, 3, 12, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = q @ k.transpose(2, 3)
    if lens is not None:
        m = torch.arange(T, device=cuda)[None, :] >= kv[:, None]
        s = s.masked_fill(m[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, 768)
    assert G.rel_err(out.float(), ref) < 1e-3    # P and O are rounded to fp16: ~2.5e-4 measured


# ------------------------------------------------------------------------------------------------ powf replay
def test_powf_half_on_device_matches_libm(lib, cuda):
    rng = np.random.default_rng(0)
    x = np.concatenate([(rng.random(300000) * 3000 + 1e-8).astype(np.float32),
                        np.float32([float.fromhex("0x1.66bf82p+8"), float.fromhex("0x1.aaf7eap+9"), 1e-8, 1.0, 768.0])])
    xd = torch.from_numpy(x).to(cuda)
    yd = torch.empty_like(xd)
    assert lib.syl_powf_half(G.ptr(xd), G.ptr(yd), x.size, G.stream()) == 0
    torch.cuda.synchronize()
    libm = ctypes.CDLL("libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    want = np.array([libm.powf(float(v), 0.5) for v in x], dtype=np.float32)
    got = yd.cpu().numpy()
    assert np.array_equal(got, want)
    assert (got != np.sqrt(x)).sum() > 0         # the replay is not just sqrt


# ------------------------------------------------------------------------------------------------ segmentation
def _check_segmentation(lib, cuda, st):
    seg, cnt, feat = G.segment(lib, torch.from_numpy(st).to(cuda))
    for b in range(st.shape[0]):
        want = R.c_get_segment(st[b], 2.6, 0.8)
        got = seg[b, :cnt[b]].astype(np.int64)
        assert len(want) == len(got) and np.array_equal(want, got), (b, want[:5], got[:5])
        if len(want):
            assert np.array_equal(R.c_segment_mean(st[b], want), feat[b, :cnt[b]], equal_nan=True)
    return cnt


@pytest.mark.parametrize("seed,B,T", [(0, 8, 499), (1, 8, 143), (2, 4, 37), (3, 2, 1500), (4, 16, 1)])
def test_segmentation_bit_exact_vs_oracle(lib, cuda, seed, B, T):
    rng = np.random.default_rng(seed)
    st = np.stack([plateau_states(rng, T) for _ in range(B)])
    cnt = _check_segmentation(lib, cuda, st)
    if T >= 100:
        assert cnt.min() >= 1 and cnt.max() < T       # a non-degenerate mix of merges and splits


def test_segmentation_edge_cases(lib, cuda):
    rng = np.random.default_rng(9)
    d = 768
    quiet = rng.standard_normal((1, 40, d)).astype(np.float32) * 0.01       # nothing above the norm threshold
    c = rng.standard_normal(d).astype(np.float32)
    c *= 3.0 / np.linalg.norm(c)
    flat = np.repeat(c[None, None], 40, 1)                                   # one segment spanning everything
    ortho = np.zeros((1, 40, d), np.float32)                                 # every frame its own segment
    for i in range(40):
        ortho[0, i, i] = 3.0
    gaps = plateau_states(rng, 40)[None].copy()
    gaps[0, ::3] *= 0.01                                                     # silence every third frame
    st = np.concatenate([quiet, flat, ortho, gaps])
    cnt = _check_segmentation(lib, cuda, st)
    assert cnt[0] == 0 and cnt[1] == 1 and cnt[2] >= 20


def _speechy(rng, T, on_mask):
    """plateau states whose frames are forced above / below the norm threshold by `on_mask`"""
    st = plateau_states(rng, T, sil=0.0)
    nrm = np.linalg.norm(st, axis=1, keepdims=True)
    st = st / nrm * rng.uniform(2.7, 3.6, (T, 1)).astype(np.float32)
    st[~on_mask] *= np.float32(0.01)
    return st.astype(np.float32)


def test_segmentation_run_structure(lib, cuda):
    """The scan is parallel over runs (stretches of frames above the norm threshold), one CTA per 16-frame chunk of run
    starts (segment.cuh): runs that start / end exactly on chunk borders, one run through every chunk, runs longer than a
    chunk behind short ones, T a multiple of 32 and not, a run reaching the last frame."""
    rng = np.random.default_rng(21)
    cases = []
    for T in (32, 33, 64, 499, 1000):
        on = np.ones(T, bool)
        cases.append((T, on.copy()))                       # one run: every CTA but the first has nothing to do
        on = np.ones(T, bool); on[15::16] = False           # every run ends on the last frame of a chunk
        cases.append((T, on.copy()))
        on = np.ones(T, bool); on[0::16] = False            # every run starts on frame 1 of a chunk
        cases.append((T, on.copy()))
        on = np.ones(T, bool); on[16::16] = False; on[0] = False
        cases.append((T, on.copy()))
        on = rng.random(T) < 0.85                           # the bench's occupancy: short runs
        cases.append((T, on.copy()))
        on = np.zeros(T, bool); on[T // 3:] = True          # silence, then one run to the end
        cases.append((T, on.copy()))
        on = np.ones(T, bool); on[5] = False; on[7] = False; on[T - 1] = False   # short runs, then a long one from chunk 0
        cases.append((T, on.copy()))
    for T, on in cases:
        st = _speechy(rng, T, on)[None]
        _check_segmentation(lib, cuda, st)


@pytest.mark.parametrize("kind", ["alternate", "tail", "drift"])
def test_segmentation_long_segments(lib, cuda, kind):
    """Refinement over LONG segments: means of 35-400 rows and sweep windows of 30-200 frames come through the
    shared-memory row ring (segment.cuh, SEG_STREAM_MIN) - same arithmetic, so still bit-identical to the oracle; "drift"
    takes the merge branch, the others the sweep."""
    rng = np.random.default_rng({"alternate": 41, "tail": 42, "drift": 43}[kind])
    for T in (300, 700, 1200):
        st = np.stack([long_segment_states(rng, T, kind) for _ in range(2)])
        cnt = _check_segmentation(lib, cuda, st)
        assert cnt.min() >= 1


def test_segmentation_reuses_workspace(lib, cuda):
    """The slot table and the finished-CTA counter live in the caller's workspace and are reset by every call: two
    different batches through ONE workspace, each equal to the oracle (a stale slot would add a segment)."""
    rng = np.random.default_rng(22)
    B, T = 4, 300
    need = lib.syl_segment_workspace_bytes(B, T)
    ws = torch.full((need,), 0x5a, dtype=torch.uint8, device=cuda)          # garbage, not zeros
    seg = torch.zeros((B, T, 2), dtype=torch.int32, device=cuda)
    cnt = torch.zeros((B,), dtype=torch.int32, device=cuda)
    for rep in range(3):
        st = np.stack([plateau_states(rng, T, sil=[0.25, 0.0, 0.6][rep]) for _ in range(B)])
        sd = torch.from_numpy(st).to(cuda)
        rc = lib.syl_segment(G.ptr(sd), B, T, float(np.float32(2.6)), float(np.float32(0.8)), G.ptr(seg), G.ptr(cnt), G.ptr(None), T,
                             G.ptr(ws), need, G.stream())
        assert rc == 0
        torch.cuda.synchronize()
        for b in range(B):
            want = R.c_get_segment(st[b], 2.6, 0.8)
            assert np.array_equal(seg[b, :int(cnt[b])].cpu().numpy().astype(np.int64), want), (rep, b)


def test_segmentation_golden(lib, cuda):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "segment_cases.npz"))
    for i in range(8):
        st = g[f"states_{i}"]
        seg, cnt, feat = G.segment(lib, torch.from_numpy(st[None]).to(cuda))
        want = g[f"segments_{i}"]
        if want.size == 0:
            assert cnt[0] == 0
        else:
            assert np.array_equal(seg[0, :cnt[0]].astype(np.int64), want)
            assert np.array_equal(feat[0, :cnt[0]], g[f"features_{i}"])
