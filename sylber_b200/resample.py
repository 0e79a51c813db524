"""Filter bank of the band-limited sinc resampler the reference applies to non-16 kHz files
(`torchaudio.transforms.Resample(sr, 16000)`, sylber/model/sylber.py:85; torchaudio's default
`sinc_interp_hann`, lowpass_filter_width 6, rolloff 0.99).  Restated from the published algorithm - one windowed-sinc
FIR per output phase, `y[i * new + p] = sum_k h[p, k] * xpad[i * orig + k]` - so that the convolution can run on the
device (`syl_resample`); checked against torchaudio's own kernel in tests/test_oracle_frontdoor.py."""
from __future__ import annotations

import math

import numpy as np


def sinc_resample_kernel(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99):
    """Returns (kernel float32 [new_g, K], width, orig_g, new_g) with K = 2 * width + orig_g, frequencies reduced by
    their gcd.  float64 arithmetic in torchaudio's operation order, cast to float32 at the end."""
    orig_freq, new_freq = int(orig_freq), int(new_freq)
    g = math.gcd(orig_freq, new_freq)
    orig_g, new_g = orig_freq // g, new_freq // g
    base = min(orig_g, new_g) * rolloff
    width = math.ceil(lowpass_filter_width * orig_g / base)
    idx = np.arange(-width, width + orig_g, dtype=np.float64)[None, :] / orig_g
    # torchaudio divides an int64 arange by new_freq, which yields float32 phases that are then promoted to float64
    phase = (np.arange(0, -new_g, -1).astype(np.float32) / np.float32(new_g)).astype(np.float64)
    t = phase[:, None] + idx
    t = t * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base / orig_g
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * (window * scale)
    return k.astype(np.float32), width, orig_g, new_g


def resampled_length(n, orig_g, new_g):
    """ceil(new * n / orig), as torchaudio truncates the convolution output."""
    return -(-new_g * int(n) // orig_g)
