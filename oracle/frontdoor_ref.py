"""TEST INFRASTRUCTURE ONLY - CPU restatements of the steps either side of the Segmenter path.

  * normalize_pcm16      sylber/model/sylber.py:83-87 (torchaudio int16 -> float = x / 32768, then (w - mean) / w.std())
                         and the zero padding of :107-111, with the same torch CPU fp32 ops the reference executes
  * kmeans_assign        sylber/model/quantizer.py:86-110 (KMQuantizer.get_indices): the codebook lookup lives in the
                         third-party vector_quantize_pytorch (requirements.txt pins 1.18.5; absent here), whose
                         EuclideanCodebook takes argmax of -cdist(x, embed); restated as an exhaustive float64 search
  * sylber_segment       sylber/model/sylber.py:208-247 (Sylber.segment) on top of oracle/segment_ref.get_segment
Parity pinning: normalize_pcm16 / sylber_segment are checked against the unmodified reference functions when
/root/reference is importable (tests/test_oracle_frontdoor.py); kmeans_assign has no reference code in the tree -
"parity unpinned" for that one function, anchored on the call site's documented semantics."""
import numpy as np
import torch

from . import segment_ref


def normalize_pcm16(pcm_list, normalize=True, sample_rate=16000):
    """list of 1-D int16 arrays -> (B, T_max) fp32 tensor, zero padded, and the lengths.  sample_rate != 16000 applies
    the reference's own resampler between the conversion and the normalisation (sylber.py:84-86)."""
    rows = []
    for x in pcm_list:
        w = torch.from_numpy(np.asarray(x, dtype=np.int16).astype(np.float32) / 32768.0)[None, :]
        if sample_rate != 16000:
            import torchaudio
            w = torchaudio.transforms.Resample(sample_rate, 16000)(w)
        if normalize:
            w = (w - w.mean()) / w.std()
        rows.append(w[0])
    lens = [int(r.shape[0]) for r in rows]
    out = torch.zeros(len(rows), max(lens))
    for i, r in enumerate(rows):
        out[i, :lens[i]] = r
    return out, lens


def resample_f64(x, sample_rate):
    """Band-limited sinc resampling to 16 kHz in float64: the reference's torchaudio.transforms.Resample (sylber.py:85)
    restated from its filter bank (sylber_b200/resample.py, bit-identical to torchaudio's - tests/test_oracle_frontdoor.py)
    and evaluated exactly; torchaudio's own fp32 conv1d is ~1e-5 away from this on unit-variance input."""
    from sylber_b200.resample import sinc_resample_kernel, resampled_length
    h, width, orig_g, new_g = sinc_resample_kernel(sample_rate, 16000)
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    K = h.shape[1]
    xpad = np.concatenate([np.zeros(width), x, np.zeros(width + orig_g)])
    win = np.lib.stride_tricks.sliding_window_view(xpad, K)[::orig_g]            # [blocks, K]
    y = (win @ h.astype(np.float64).T).reshape(-1)                               # [blocks * new_g]
    return y[:resampled_length(len(x), orig_g, new_g)]


def kmeans_assign(feats, centroids, normalize=False):
    x = np.asarray(feats, dtype=np.float64).reshape(-1, centroids.shape[1])
    if normalize:
        x32 = np.asarray(feats, dtype=np.float32).reshape(-1, centroids.shape[1])
        x = (x32 / (((x32 ** 2).sum(-1) + 1e-8) ** .5)[..., None] * 6).astype(np.float64)
    c = np.asarray(centroids, dtype=np.float64)
    d = (x * x).sum(-1)[:, None] - 2.0 * x @ c.T + (c * c).sum(-1)[None, :]
    return d.argmin(-1), d


def sylber_segment(features, normthreshold, mergethreshold):
    """features (B,T,768) fp32 array -> (segments list, avg_fts (B, max(N,1), 768) zero padded), sylber.py:219-245."""
    segments = [segment_ref.get_segment(states, normthreshold, mergethreshold) for states in features]
    n_max = max(max((len(s) for s in segments), default=0), 1)
    avg = np.zeros((len(features), n_max, features.shape[-1]), dtype=np.float32)
    for b, segs in enumerate(segments):
        for j, (s, e) in enumerate(np.asarray(segs).reshape(-1, 2).astype(np.int64)):
            avg[b, j] = torch.from_numpy(features[b][s:e]).mean(0).numpy()     # torch mean, as the reference
    return segments, avg
