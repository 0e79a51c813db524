"""Norm-threshold estimator of the reference (sylber/utils/segment_utils.py:6-59, `Thresholder`): the decision boundary
between two Gaussians fitted to the frame norms of speech ("signal") and non-speech ("noise") frames, tracked with
an exponential moving average.  Host-side float32 arithmetic in the reference's operation order; no torch modules,
so a stage-2 checkpoint's four statistics can drive `Segmenter.segment(normthreshold=None)`."""
from __future__ import annotations

import numpy as np

_F = np.float32


class Thresholder:
    def __init__(self, signal_mean=None, signal_var=None, noise_mean=None, noise_var=None, decay=0.9999, eta=1,
                 threshold=None):
        fixed = threshold is not None
        self.signal_mean = _F(1.0 if fixed else signal_mean)
        self.signal_var = _F(1.0 if fixed else signal_var)
        self.noise_mean = _F(1.0 if fixed else noise_mean)
        self.noise_var = _F(1.0 if fixed else noise_var)
        self.decay = decay
        self.eta = eta
        self.threshold = _F(threshold) if fixed else None

    @classmethod
    def from_state_dict(cls, sd, prefix="thresholder.", **kw):
        """Statistics as saved by the reference's nn.Module (keys signal_mean, signal_var, noise_mean, noise_var)."""
        get = lambda k: float(np.asarray(sd[prefix + k]).reshape(-1)[0])
        if prefix + "threshold" in sd:
            return cls(threshold=get("threshold"), **kw)
        return cls(get("signal_mean"), get("signal_var"), get("noise_mean"), get("noise_var"), **kw)

    def get_threshold(self):
        """segment_utils.py:27-52.  Returns a float32 scalar (None if the quadratic has no real root, as the
        reference's unbound local would raise)."""
        if self.threshold is not None:
            return self.threshold
        mu_s = self.signal_mean
        sigma_s = _F((self.signal_var + _F(1e-8)) ** _F(0.5))
        mu_n = self.noise_mean
        sigma_n = _F((self.noise_var + _F(1e-8)) ** _F(0.5))
        a = _F(sigma_s ** 2 - sigma_n ** 2)
        b = _F(_F(-2) * sigma_s ** 2 * mu_n + _F(2) * sigma_n ** 2 * mu_s)
        log_term = _F(_F(np.log(self.eta)) + _F(np.log(_F(sigma_s / sigma_n))))
        c = _F(sigma_s ** 2 * mu_n ** 2 - sigma_n ** 2 * mu_s ** 2 - _F(2) * sigma_n ** 2 * sigma_s ** 2 * log_term)
        if a != 0:
            disc = _F(b ** 2 - _F(4) * a * c)
            if disc > 0:
                sign = _F(1.0) if mu_s > mu_n else _F(0.0)
                return _F((-b + sign * _F(np.sqrt(disc))) / (_F(2) * a))
            if disc == 0:
                return _F(-b / (_F(2) * a))
            return None
        if b != 0:
            return _F(-c / b)
        return None

    def update_stats(self, signal, noise):
        """segment_utils.py:55-65: EMA of mean and (biased, around the NEW mean) variance of the two populations."""
        if self.threshold is not None:
            return
        d, e = _F(self.decay), _F(1 - self.decay)
        if signal is not None and len(signal):
            s = np.asarray(signal, dtype=np.float32)
            self.signal_mean = _F(d * self.signal_mean + e * s.mean(dtype=np.float32))
            self.signal_var = _F(d * self.signal_var + e * ((s - self.signal_mean) ** 2).mean(dtype=np.float32))
        if noise is not None and len(noise):
            n = np.asarray(noise, dtype=np.float32)
            self.noise_mean = _F(d * self.noise_mean + e * n.mean(dtype=np.float32))
            self.noise_var = _F(d * self.noise_var + e * ((n - self.noise_mean) ** 2).mean(dtype=np.float32))
