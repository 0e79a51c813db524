"""Host-side mirror of the reference's inference API (sylber/model/sylber.py:28-138).

`Segmenter` keeps the reference's constructor, call signature, attributes and output contract; what changes
is that lines 122-133 of the reference (HubertModel forward, hidden_states.cpu(), get_segment, per-segment
means) become ONE call into the C ABI (`syl_forward`), which runs the conv front end, the transformer
encoder, the segmentation scan and the pooling on the GPU and hands back NumPy arrays.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .weights import normalize_state_dict, random_hubert_state_dict, REQUIRED_KEYS

HIDDEN = 768
FRAME_RATE = 50


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _Engine:
    """Owns the C-ABI handle, the device workspace and the launch of syl_forward."""

    def __init__(self, state_dict, n_layers, device, mode):
        if not torch.cuda.is_available():
            raise RuntimeError("sylber_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"sylber_b200 runs on CUDA devices only, got {device!r}")
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.index)
        self.n_layers = n_layers
        self.mode = _lib.MODES[mode] if isinstance(mode, str) else int(mode)
        handle = ctypes.c_void_p()
        rc = self.lib.syl_create(ctypes.byref(handle), self.index, n_layers, self.mode)
        _lib.check(self.lib, None, rc, "syl_create")
        self.handle = handle
        self._workspace = None
        missing = [k for k in REQUIRED_KEYS(n_layers) if k not in state_dict]
        if missing:
            raise RuntimeError(f"checkpoint is missing {len(missing)} tensors, first: {missing[:4]}")
        with torch.cuda.device(self.index):
            for name in REQUIRED_KEYS(n_layers):
                t = state_dict[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                rc = self.lib.syl_load_weight(self.handle, name.encode(), _ptr(t), shape, t.dim(), 0)
                _lib.check(self.lib, self.handle, rc, f"syl_load_weight({name})")
            torch.cuda.synchronize(self.index)
            rc = self.lib.syl_finalize(self.handle)
            _lib.check(self.lib, self.handle, rc, "syl_finalize")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.syl_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def num_frames(self, n_samples):
        return int(self.lib.syl_num_frames(int(n_samples)))

    def workspace(self, batch, t_samp):
        need = int(self.lib.syl_workspace_bytes(self.handle, batch, t_samp))
        if need == 0:
            raise ValueError(f"invalid shape batch={batch} samples={t_samp} (need >= 400 samples)")
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace, need

    def forward(self, wav, n_samples, thr_norm, thr_merge, segment=True, max_seg=None):
        """wav (B, T_samp) fp32 cuda, n_samples (B,) int32 cuda or None.

        Returns hidden (B,T,768) and, if `segment`, (seg (B,max_seg,2) int32, seg_count (B,) int32,
        seg_feat (B,max_seg,768) fp32), all device tensors on the current stream."""
        B, t_samp = wav.shape
        T = self.num_frames(t_samp)
        ws, need = self.workspace(B, t_samp)
        hidden = torch.empty((B, T, HIDDEN), dtype=torch.float32, device=self.device)
        seg = cnt = feat = None
        if segment:
            max_seg = T if max_seg is None else int(max_seg)
            seg = torch.empty((B, max_seg, 2), dtype=torch.int32, device=self.device)
            cnt = torch.empty((B,), dtype=torch.int32, device=self.device)
            feat = torch.empty((B, max_seg, HIDDEN), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.syl_forward(self.handle, _ptr(wav), _ptr(n_samples), B, t_samp, _ptr(hidden), _ptr(seg), _ptr(cnt),
                                  _ptr(feat), max_seg if segment else 0, float(thr_norm), float(thr_merge),
                                  _ptr(ws), need, ctypes.c_void_p(stream))
        _lib.check(self.lib, self.handle, rc, "syl_forward")
        return hidden, seg, cnt, feat

    def read_stage(self, name, shape):
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.syl_read_stage(self.handle, name.encode(), _ptr(out), out.numel(), ctypes.c_void_p(stream))
        _lib.check(self.lib, self.handle, rc, f"syl_read_stage({name})")
        return out

    def set_active_layers(self, n):
        _lib.check(self.lib, self.handle, self.lib.syl_set_active_layers(self.handle, int(n)), "syl_set_active_layers")

    def profile(self, on):
        _lib.check(self.lib, self.handle, self.lib.syl_profile_enable(self.handle, int(bool(on))), "syl_profile_enable")

    def profile_read(self):
        """{stage name: (milliseconds, regions)} accumulated since the previous read."""
        n = self.lib.syl_num_stages()
        ms = (ctypes.c_float * n)()
        cnt = (ctypes.c_int * n)()
        _lib.check(self.lib, self.handle, self.lib.syl_profile_read(self.handle, ms, cnt), "syl_profile_read")
        return {self.lib.syl_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def launch_count(self, with_segmentation=True):
        return int(self.lib.syl_forward_launch_count(self.handle, int(with_segmentation)))


class SpeechModel:
    """What the reference keeps in `Segmenter.speech_model` (a HubertModel): callable as
    `(input_values, attention_mask=None, mask_time_indices=None)` returning `.last_hidden_state`
    (sylber/model/sylber.py:122,214,257).  Backed by the CUDA engine, inference only."""

    def __init__(self, engine):
        self._engine = engine
        self.config = SimpleNamespace(hidden_size=HIDDEN, num_hidden_layers=engine.n_layers, num_attention_heads=12,
                                      intermediate_size=3072)

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    @torch.no_grad()
    def __call__(self, input_values, attention_mask=None, mask_time_indices=None, **_kw):
        if mask_time_indices is not None:
            raise NotImplementedError("mask_time_indices is a training-time feature; sylber_b200 is inference only")
        eng = self._engine
        wav = input_values.to(device=eng.device, dtype=torch.float32).contiguous()
        n = None
        if attention_mask is not None:
            n = attention_mask.to(eng.device).sum(-1).to(torch.int32).contiguous()
        hidden, _, _, _ = eng.forward(wav, n, 0.0, 0.0, segment=False)
        return SimpleNamespace(last_hidden_state=hidden)

    forward = __call__


def _read_audio(path):
    """(channels, samples) float32 in [-1, 1) and the sample rate.  torchaudio when it works, else a WAV reader."""
    try:
        import torchaudio  # noqa: WPS433
        wav, sr = torchaudio.load(path)
        return wav.to(torch.float32), int(sr)
    except Exception:
        from scipy.io import wavfile  # noqa: WPS433
        sr, data = wavfile.read(path)
        if data.dtype == np.int16:
            x = data.astype(np.float32) / 32768.0
        elif data.dtype == np.int32:
            x = data.astype(np.float32) / 2147483648.0
        elif data.dtype == np.uint8:
            x = (data.astype(np.float32) - 128.0) / 128.0
        else:
            x = data.astype(np.float32)
        x = x[None, :] if x.ndim == 1 else x.T
        return torch.from_numpy(np.ascontiguousarray(x)), int(sr)


def _resample(wav, sr):
    import torchaudio  # noqa: WPS433
    return torchaudio.transforms.Resample(sr, 16000)(wav)


class Segmenter:
    """B200-native drop-in for `sylber.Segmenter` (sylber/model/sylber.py:28-138).

    Differences from the reference, all opt-in or forced by the hardware target:
      * `device` must be a CUDA device; a missing GPU raises instead of silently falling back to CPU
        (the reference's fallback at :56-58 runs after `.to(device)` has already raised).
      * missing checkpoint tensors raise (the reference's strict=False at :52 ignores them).
      * extra keyword arguments: `state_dict=` (use these tensors instead of loading `model_ckpt`),
        `mode=` ("parity" default | "fast" | "exact", see include/sylber_b200.h), `max_batch=`.
    """

    def __init__(self,
                 model_ckpt="sylber",
                 speech_upstream="facebook/hubert-base-ls960",
                 ema_decay=0.999,
                 encoding_layer=9,
                 merge_threshold=0.8,
                 norm_threshold=2.6,
                 device='cuda',
                 **kwargs):
        state_dict = kwargs.pop("state_dict", None)
        mode = kwargs.pop("mode", "parity")
        self.max_batch = int(kwargs.pop("max_batch", 64))
        self.enc_dim = HIDDEN
        self.encoding_layer = encoding_layer
        self.ema_decay = ema_decay
        self.speech_upstream = speech_upstream

        if state_dict is None:
            if model_ckpt is not None:
                if model_ckpt == "sylber":
                    model_ckpt = "sylber.ckpt"
                if not Path(model_ckpt).exists():
                    from huggingface_hub import hf_hub_download  # same source as sylber.py:49-50
                    model_ckpt = hf_hub_download(repo_id="cheoljun95/sylber", filename=model_ckpt)
                state_dict = torch.load(model_ckpt, map_location="cpu")
                print("Pre-trained checkpoint loaded")
            else:
                # the reference leaves HubertModel randomly initialised when model_ckpt is None (sylber.py:41-46)
                state_dict = random_hubert_state_dict(encoding_layer)
        state_dict = normalize_state_dict(state_dict)

        if 'cuda' in str(device) and not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available and sylber_b200 has no CPU path")
        self._engine = _Engine(state_dict, encoding_layer, device, mode)
        self.speech_model = SpeechModel(self._engine)
        self.device = str(self._engine.device)
        self.norm_threshold = norm_threshold
        self.merge_threshold = merge_threshold

    # ------------------------------------------------------------------------------------------
    def _prepare(self, wav_file, wav):
        """sylber.py:76-101: load / normalise / collect lengths.  Returns (list of (1,T) fp32 cpu tensors, is_batch)."""
        batch_wavs = []
        if wav_file is not None:
            is_batch = isinstance(wav_file, list)
            wav_files = wav_file if is_batch else [wav_file]
            for file in wav_files:
                w, sr = _read_audio(os.fspath(file))
                if sr != 16000:
                    w = _resample(w, sr)
                w = (w - w.mean()) / w.std()
                batch_wavs.append(w)
        else:
            assert wav is not None
            is_batch = isinstance(wav, list)
            batch_wavs = wav if is_batch else [wav]
        return batch_wavs, is_batch

    @torch.no_grad()
    def __call__(self, wav_file=None, wav=None, in_second=True):
        """Same contract as the reference: a dict (single input) or list of dicts with
        `segments` (N,2), `segment_features` (N,768) float32, `hidden_states` (T_max,768) float32."""
        batch_wavs, is_batch = self._prepare(wav_file, wav)
        eng = self._engine
        rows = []
        for w in batch_wavs:
            w = torch.as_tensor(w)
            if w.dim() == 1:
                w = w[None, :]
            rows.extend(w[i] for i in range(w.shape[0]))      # torch.cat(dim=0) at sylber.py:117: channels become rows
        lengths = [int(r.shape[-1]) for r in rows]
        max_length = max(lengths)
        outputs = []
        for lo in range(0, len(rows), self.max_batch):
            chunk = rows[lo:lo + self.max_batch]
            # pinned staging (torch's caching host allocator makes these cheap after the first call)
            host = torch.empty((len(chunk), max_length), dtype=torch.float32, pin_memory=True)
            for i, r in enumerate(chunk):
                n = r.shape[-1]
                host[i, :n].copy_(r)
                if n < max_length:
                    host[i, n:].zero_()
            n_host = torch.tensor(lengths[lo:lo + len(chunk)], dtype=torch.int32).pin_memory()
            wav_dev = host.to(eng.device, non_blocking=True)
            n_dev = n_host.to(eng.device, non_blocking=True)
            hidden, seg, cnt, feat = eng.forward(wav_dev, n_dev, np.float32(self.norm_threshold),
                                                 np.float32(self.merge_threshold))
            hidden_pin = torch.empty(hidden.shape, dtype=torch.float32, pin_memory=True)
            hidden_pin.copy_(hidden, non_blocking=True)
            cnt_h = cnt.cpu().numpy()                      # synchronises the stream
            n_max = max(int(cnt_h.max()) if len(cnt_h) else 0, 1)
            seg_pin = torch.empty((len(chunk), n_max, 2), dtype=torch.int32, pin_memory=True)
            feat_pin = torch.empty((len(chunk), n_max, HIDDEN), dtype=torch.float32, pin_memory=True)
            seg_pin.copy_(seg[:, :n_max], non_blocking=True)
            feat_pin.copy_(feat[:, :n_max], non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
            hidden_h, seg_h, feat_h = hidden_pin.numpy(), seg_pin.numpy(), feat_pin.numpy()
            for i in range(len(chunk)):
                n = int(cnt_h[i])
                segments = seg_h[i, :n].astype(np.int64) if n > 0 else np.array([])
                outputs.append({
                    'segments': segments * 1.0 / FRAME_RATE if in_second else segments,
                    'segment_features': feat_h[i, :n] if n > 0 else np.array([]),
                    'hidden_states': hidden_h[i],
                })
        return outputs if is_batch else outputs[0]
