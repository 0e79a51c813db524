"""CPU oracle (TEST INFRASTRUCTURE): end-to-end segment agreement between a device run and the fp32 CPU path.

`get_segment` (sylber/utils/segment_utils.py:72-131) is a discontinuous function of the hidden states: `norms >=
normthreshold` (:76), `sim >= mergethreshold` (:96-97, :114) and the argmax at :126 change outcome on arbitrarily
small state differences, so "identical segments" between two implementations of the forward holds only for
utterances none of whose decisions sits closer to its threshold than the state error.  This module measures that:
for one utterance it compares the segments computed from the oracle's fp32 states with the ones the device produced,
and when they differ it names the first decision that flipped, its margin in the oracle's run and how far the value
moved, so that every disagreement is either explained by margin <= state error or is a bug.

Only tests/, __graft_entry__.smoke() and bench.py's cpu/reference legs may import this module.
"""
from __future__ import annotations

import numpy as np

from . import segment_ref as R


def same_segments(a, b):
    a = np.asarray(a).reshape(-1, 2).astype(np.int64)
    b = np.asarray(b).reshape(-1, 2).astype(np.int64)
    return a.shape == b.shape and bool(np.array_equal(a, b))


def frame_errors(ref_states, got_states):
    """(relative Frobenius error, largest per-frame relative error) of got vs ref, both (T, d)."""
    r = np.asarray(ref_states, np.float64)
    d = np.asarray(got_states, np.float64) - r
    per = np.linalg.norm(d, axis=1) / np.maximum(np.linalg.norm(r, axis=1), 1e-30)
    return float(np.linalg.norm(d) / np.linalg.norm(r)), float(per.max())


def flip_is_explained(flip, max_frame_err):
    """Is the first flipped decision within what a per-frame relative state error of `max_frame_err` can do?
    |d norm| / norm <= err;  |d cos(u, v)| <= 2 (err_u + err_v) <= 4 err (centroids are means of frames, so their
    error is at most the frames');  a boundary sweep (:125) sums at most `window` cosines, each moving by <= 4 err,
    and a window is at most 64 frames wide in practice, so a lead below 2 * 64 * 4 err can change hands."""
    if flip is None:
        return True
    if flip["kind"] == "norm":
        return flip["delta"] <= 2.0 * max_frame_err * abs(flip["ref"])
    if flip["kind"] in ("merge", "refine_merge"):
        return flip["delta"] <= 4.0 * max_frame_err
    return flip["margin"] <= 512.0 * max_frame_err


def compare_utterance(ref_states, got_states, got_segments, norm_thr=2.6, merge_thr=0.8):
    """One utterance: oracle states (T,768) vs the device's states and the device's segments.  Returns a dict with
    agree, n_ref, n_got, rel, max_frame_err, margins (oracle run) and, on disagreement, first_flip + explained."""
    ref_states = np.ascontiguousarray(ref_states, np.float32)
    got_states = np.ascontiguousarray(got_states, np.float32)
    ref_seg = R.c_get_segment(ref_states, norm_thr, merge_thr)
    rel, worst = frame_errors(ref_states, got_states)
    rec = {"agree": same_segments(ref_seg, got_segments), "n_ref": int(len(ref_seg)),
           "n_got": int(len(np.asarray(got_segments).reshape(-1, 2))), "rel": rel, "max_frame_err": worst}
    _, trace = R.decision_trace(ref_states, norm_thr, merge_thr)
    rec["margins"] = R.min_margins(trace)
    rec["decisions"] = len(trace)
    if not rec["agree"]:
        flip = R.explain_difference(ref_states, got_states, norm_thr, merge_thr)
        rec["first_flip"] = flip
        rec["explained"] = flip is not None and flip_is_explained(flip, worst)
    return rec


def summarize(records):
    """Batch summary: agreement count, worst errors, smallest margins, and the flips."""
    n = len(records)
    out = {"utterances": n, "agree": sum(r["agree"] for r in records),
           "rel_max": max(r["rel"] for r in records), "max_frame_err": max(r["max_frame_err"] for r in records),
           "decisions": sum(r["decisions"] for r in records),
           "min_margin": {k: min(r["margins"].get(k, float("inf")) for r in records)
                          for k in ("norm", "merge", "refine_merge", "refine_argmax")},
           "flips": [dict(r["first_flip"] or {"kind": "none"}, utterance=i, explained=r["explained"])
                     for i, r in enumerate(records) if not r["agree"]]}
    out["all_flips_explained"] = all(f["explained"] for f in out["flips"])
    return out
