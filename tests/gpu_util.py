"""Helpers shared by the GPU parity tests: thin ctypes callers for the per-stage C-ABI entry points."""
import ctypes

import torch


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel_err(a, b):
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def gemm_f32(lib, A, W, bias=None, residual=None, n_pass=1, act=0):
    M, K = A.shape
    N = W.shape[0]
    need = lib.syl_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=A.device)
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    rc = lib.syl_gemm_f32(ptr(A), ptr(W), ptr(bias), ptr(residual), ptr(out), M, N, K, n_pass, act, ptr(ws), need, stream())
    if rc != 0:
        raise RuntimeError(f"syl_gemm_f32 rc={rc}: {lib.syl_last_error(None)}")
    torch.cuda.synchronize()
    return out


def segment(lib, states, thr_norm=2.6, thr_merge=0.8, with_feat=True):
    import numpy as np
    B, T, D = states.shape
    need = lib.syl_segment_workspace_bytes(B, T)
    ws = torch.empty(need, dtype=torch.uint8, device=states.device)
    seg = torch.zeros((B, T, 2), dtype=torch.int32, device=states.device)
    cnt = torch.zeros((B,), dtype=torch.int32, device=states.device)
    feat = torch.zeros((B, T, D), dtype=torch.float32, device=states.device) if with_feat else None
    rc = lib.syl_segment(ptr(states), B, T, float(np.float32(thr_norm)), float(np.float32(thr_merge)), ptr(seg), ptr(cnt),
                         ptr(feat), T, ptr(ws), need, stream())
    if rc != 0:
        raise RuntimeError(f"syl_segment rc={rc}")
    torch.cuda.synchronize()
    return seg.cpu().numpy(), cnt.cpu().numpy(), (feat.cpu().numpy() if with_feat else None)
