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
