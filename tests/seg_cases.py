"""Synthetic hidden-state generators for segmentation parity tests (shared by CPU and GPU tests)."""
import numpy as np


def plateau_states(rng, T, d=768, noise=0.35, sil=0.25):
    """Syllable-like states: piecewise-constant directions plus noise, low-norm 'silence' gaps, and
    neighbouring plateaus that are deliberately close to the merge threshold."""
    out = np.zeros((T, d), np.float32)
    t = 0
    while t < T:
        L = int(rng.integers(1, 14))
        n = min(L, T - t)
        if rng.random() < sil:
            out[t:t + n] = rng.standard_normal((n, d)).astype(np.float32) * 0.05
        else:
            c = rng.standard_normal(d).astype(np.float32)
            c *= rng.uniform(2.4, 4.0) / np.linalg.norm(c)
            if rng.random() < 0.4 and t > 0:
                c = 0.8 * out[t - 1] + 0.6 * c * rng.uniform(0.3, 1.2)
            z = rng.standard_normal((n, d)).astype(np.float32)
            z *= noise * np.linalg.norm(c) / np.sqrt(d) * rng.uniform(0.2, 2.5)
            out[t:t + n] = c[None] + z
        t += L
    return out


def long_segment_states(rng, T, kind, d=768):
    """States whose segments are LONG (35-160 frames) on both sides of the mid-boundaries, so that the refinement phase
    (segment_utils.py:110-128) works on means and sweep windows of 30-400 rows - the shape the padded tails of short clips
    produce.  kind: "alternate" (two directions taking turns: every boundary is refined by the sweep), "tail" (one
    direction with single outlier frames that split it), "drift" (neighbouring plateaus at cosine 0.80-0.92 and noisy
    frames: the scan splits, the segment means merge again)."""
    def unit(v):
        return v / np.linalg.norm(v)
    out = np.zeros((T, d), np.float32)
    c1 = unit(rng.standard_normal(d))
    c2 = unit(0.3 * c1 + rng.standard_normal(d))
    c = c1
    t = k = 0
    while t < T:
        n = int(rng.integers(35, 160))
        m = min(n, T - t)
        if kind == "alternate":
            c = c1 if k % 2 == 0 else c2
        r = 0.4 if kind == "drift" else 0.25
        blk = (c[None] + rng.standard_normal((m, d)) * (r / np.sqrt(d))) * 3.0
        if kind == "tail" and m > 10:
            j = int(rng.integers(3, m - 3))
            blk[j] = unit(c1 * 0.6 + unit(rng.standard_normal(d)) * 0.8) * 3.0
        out[t:t + m] = blk
        t += n
        k += 1
        if kind == "drift":
            u = rng.standard_normal(d)
            u -= (u @ c) * c
            cs = rng.uniform(0.80, 0.92)
            c = cs * c + np.sqrt(1 - cs * cs) * unit(u)
    return out.astype(np.float32)
