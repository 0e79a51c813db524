"""CPU oracle (TEST INFRASTRUCTURE): NumPy restatement of the reference's segmentation and pooling.

Follows sylber/utils/segment_utils.py:68-131 (cossim, get_segment) and the output packaging at
sylber/model/sylber.py:128-136 step for step, using the same NumPy float32 operations, so it is
bit-identical to the reference by construction and costs the same per utterance (it is what the
`cpu_baseline` / `--impl reference` legs of bench.py time, because /root/reference does not exist
on the GPU box).  `segment_oracle.c` is the fast C twin used for large randomized parity sweeps;
`c_get_segment` below is its ctypes binding.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_EPS = 1e-8


def _cos(u, v):
    # segment_utils.py:68-69 - note the epsilon sits inside each square root
    num = (u * v).sum(-1)
    den_u = ((u ** 2).sum(-1) + _EPS) ** .5
    den_v = ((v ** 2).sum(-1) + _EPS) ** .5
    return num / den_u / den_v


def _scan(states, active, merge_thr, trace=None):
    """Phase 1 (segment_utils.py:79-108): greedy left-to-right merge of frames into runs.
    `trace` (a list) receives every threshold decision as (kind, frame, value, threshold)."""
    runs, splits = [], []
    start, centroid, count = -1, 0, 0
    for t in range(len(states)):
        if not active[t]:
            if start > -1:
                runs.append([start, t])
            start, centroid, count = -1, 0, 0
            continue
        if count == 0:
            centroid, count, start = states[t], 1, t
            continue
        sim = _cos(centroid, states[t])
        if trace is not None:
            trace.append(("merge", t, float(sim), float(np.float32(merge_thr))))
        if sim >= merge_thr:
            centroid = (centroid * count + states[t]) / (count + 1)
            count += 1
        else:
            centroid = states[t]
            count += 1          # not reset: reference quirk at segment_utils.py:103
            runs.append([start, t])
            splits.append((t, len(runs) - 1))
            start = t
    if start > -1:
        runs.append([start, len(states)])
    return runs, splits


def _refine(states, runs, splits, merge_thr, trace=None):
    """Phase 2 (segment_utils.py:110-128): merge or re-place each split boundary, in order.
    `trace` receives ("refine_merge", cut, cos, thr) and ("refine_argmax", cut, best - runner_up, winning frame)."""
    absorbed = set()
    for cut, left in splits:
        if left >= len(runs) - 1:
            continue
        right = left + 1
        (ls, le), (rs, re) = runs[left], runs[right]
        mu_l = states[ls:le].mean(0)
        mu_r = states[rs:re].mean(0)
        sim = _cos(mu_l, mu_r)
        if trace is not None:
            trace.append(("refine_merge", cut, float(sim), float(np.float32(merge_thr))))
        if sim >= merge_thr:
            runs[right] = [ls, re]
            absorbed.add(left)
            continue
        lo = max(ls, cut - max(1, (le - ls) // 2))
        hi = min(re, cut + max(1, (re - rs) // 2))
        window = states[lo:hi]
        to_left = _cos(window, mu_l[None, :])
        to_right = _cos(window, mu_r[None, :])
        score = [to_left[:k].sum() + to_right[k:].sum() for k in range(hi - lo)]
        best = lo + int(np.argmax(score))
        if trace is not None:
            top = np.sort(np.asarray(score, np.float64))[::-1]
            trace.append(("refine_argmax", cut, float(top[0] - top[1]) if len(top) > 1 else float("inf"), best))
        runs[left] = [ls, best]
        runs[right] = [best, re]
    return [r for k, r in enumerate(runs) if k not in absorbed]


def get_segment(states, norm_thr, merge_thr, norms=None, trace=None):
    """Same contract as the reference's get_segment: (N,2) int64, or shape (0,) float64 when empty.
    With `trace` (a list) every threshold decision the run takes is appended in order (decision_trace below)."""
    if norms is None:
        norms = ((states ** 2).sum(-1) + _EPS) ** .5
    if trace is not None:
        thr = float(np.float32(norm_thr))
        trace.extend(("norm", t, float(v), thr) for t, v in enumerate(norms))
    runs, splits = _scan(states, norms >= norm_thr, merge_thr, trace)
    return np.array(_refine(states, runs, splits, merge_thr, trace))


def decision_trace(states, norm_thr, merge_thr):
    """(segments, decisions): every comparison get_segment makes on `states`, in execution order, as
    (kind, frame, value, threshold) with kind in norm | merge | refine_merge (segment_utils.py:76, :96-97, :114), and
    ("refine_argmax", cut, lead of the winning boundary over the runner-up, winning frame) for :126."""
    trace = []
    seg = get_segment(np.asarray(states, np.float32), norm_thr, merge_thr, trace=trace)
    return seg, trace


def min_margins(trace):
    """Smallest distance to a different outcome per decision kind: |value - threshold| / threshold for norm,
    |value - threshold| for the cosines, the winner's lead for the argmax."""
    out = {}
    for kind, _, v, thr in trace:
        m = v if kind == "refine_argmax" else (abs(v - thr) / thr if kind == "norm" else abs(v - thr))
        out[kind] = min(out.get(kind, float("inf")), m)
    return out


def _outcome(d):
    return d[3] if d[0] == "refine_argmax" else d[2] >= d[3]


def explain_difference(states_ref, states_got, norm_thr, merge_thr):
    """Why do two state arrays of one utterance segment differently?  Replays get_segment on both and returns the
    FIRST decision whose outcome differs: dict(kind, frame, ref, got, threshold, margin, delta).  For the threshold
    decisions margin = |ref - threshold| and delta = |got - ref|, so a flip always has margin <= delta and what the
    caller has to check is that delta is no larger than the per-frame state error allows.  For refine_argmax ref / got
    are the two winners' leads, margin = the reference winner's lead, delta = the sum of both leads (how far the score
    differences moved).  None when both runs take identical decisions."""
    _, a = decision_trace(states_ref, norm_thr, merge_thr)
    _, b = decision_trace(states_got, norm_thr, merge_thr)
    for da, db in zip(a, b):
        assert da[:2] == db[:2], (da, db)          # identical outcomes so far => identical control flow
        if _outcome(da) != _outcome(db):
            if da[0] == "refine_argmax":
                return {"kind": da[0], "frame": da[1], "ref": da[2], "got": db[2], "threshold": 0.0,
                        "margin": da[2], "delta": da[2] + db[2]}
            return {"kind": da[0], "frame": da[1], "ref": da[2], "got": db[2], "threshold": da[3],
                    "margin": abs(da[2] - da[3]), "delta": abs(db[2] - da[2])}
    return None


def package(states, segments, in_second=True):
    """sylber/model/sylber.py:130-135 for one utterance."""
    return {
        "segments": segments * 1.0 / 50 if in_second else segments,
        "segment_features": (np.stack([states[s:e].mean(0) for s, e in segments])
                             if len(segments) > 0 else np.array([])),
        "hidden_states": states,
    }


# --------------------------------------------------------------------------------------------
# ctypes binding of the C twin
# --------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force=False):
    so = os.path.join(_HERE, "liboracle_segment.so")
    src = os.path.join(_HERE, "segment_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_segment.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        lib.syl_oracle_get_segment.restype = ctypes.c_int64
        lib.syl_oracle_get_segment.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, ctypes.c_float, fp, ip]
        lib.syl_oracle_segment_mean.restype = None
        lib.syl_oracle_segment_mean.argtypes = [fp, ctypes.c_int64, ip, ctypes.c_int64, fp]
        lib.syl_oracle_np_sum.restype = ctypes.c_float
        lib.syl_oracle_np_sum.argtypes = [fp, ctypes.c_int64]
        lib.syl_oracle_powf_half.restype = ctypes.c_float
        lib.syl_oracle_powf_half.argtypes = [ctypes.c_float]
        _LIB = lib
    return _LIB


def c_get_segment(states, norm_thr, merge_thr):
    """C oracle: returns (N,2) int64 (N may be 0)."""
    st = np.ascontiguousarray(states, dtype=np.float32)
    T, d = st.shape
    out = np.zeros((max(T, 1), 2), dtype=np.int64)
    fp = ctypes.POINTER(ctypes.c_float)
    n = _lib().syl_oracle_get_segment(st.ctypes.data_as(fp), T, d, float(np.float32(norm_thr)),
                                      float(np.float32(merge_thr)), None,
                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return out[:n].copy()


def c_segment_mean(states, segments):
    st = np.ascontiguousarray(states, dtype=np.float32)
    seg = np.ascontiguousarray(segments, dtype=np.int64).reshape(-1, 2)
    out = np.zeros((len(seg), st.shape[1]), dtype=np.float32)
    if len(seg):
        _lib().syl_oracle_segment_mean(st.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), st.shape[1],
                                       seg.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(seg),
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def c_np_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(_lib().syl_oracle_np_sum(a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), len(a)))
