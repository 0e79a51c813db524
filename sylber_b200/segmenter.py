"""Host-side mirror of the reference's inference API (sylber/model/sylber.py:28-138).

`Segmenter` keeps the reference's constructor, call signature, attributes and output contract; what changes
is that lines 122-133 of the reference (HubertModel forward, hidden_states.cpu(), get_segment, per-segment
means) become ONE call into the C ABI (`syl_forward`), which runs the conv front end, the transformer
encoder, the segmentation scan and the pooling on the GPU and hands back NumPy arrays.
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .batching import plan_length_buckets, sub_batch_bounds
from .resample import sinc_resample_kernel, resampled_length
from .weights import normalize_state_dict, random_hubert_state_dict, REQUIRED_KEYS

HIDDEN = 768
FRAME_RATE = 50


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _PinnedPool:
    """Pinned host buffers for results.  cudaHostAlloc costs ~3 ms per call for the 49 MB hidden-state block
    (measured, tools/e2e_breakdown.py), so blocks are recycled - but only once nothing outside the pool can still
    see them: every array handed to the caller is a NumPy view whose base chain ends in one root ndarray per
    block, the pool keeps only a weak reference to that root, and a block is reused when the root has died."""

    def __init__(self):
        self._blocks = []          # [tensor, weakref-to-root-ndarray or None]

    def _make(self, n_bytes):
        return torch.empty(max(int(n_bytes), 1), dtype=torch.uint8, pin_memory=True)

    def array(self, shape, dtype=np.float32):
        """A NumPy array of `shape` backed by pinned memory, plus the torch view to copy into."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        slot = None
        for ent in self._blocks:
            if ent[0].numel() >= n and (ent[1] is None or ent[1]() is None):
                slot = ent
                break
        if slot is None:
            self._blocks = [e for e in self._blocks if e[1] is not None and e[1]() is not None][-6:]   # drop idle blocks
            slot = [self._make(n), None]
            self._blocks.append(slot)
        root = slot[0].numpy()
        slot[1] = weakref.ref(root)
        arr = root[:n].view(dtype).reshape(shape)
        ten = slot[0][:n].view(torch.from_numpy(np.empty(0, dtype)).dtype).view(*shape)
        return arr, ten


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
        self._workspaces = {}
        self._io = {}
        self._seg_io = {}
        self._stage_in = {}
        self._streams = []
        self._copy_streams = []
        self._resamplers = {}
        self.pool = _PinnedPool()
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

    def workspace(self, batch, t_samp, slot=0):
        """Device workspace for one in-flight forward; `slot` separates concurrent sub-batches."""
        need = int(self.lib.syl_workspace_bytes(self.handle, batch, t_samp))
        if need == 0:
            raise ValueError(f"invalid shape batch={batch} samples={t_samp} (need >= 400 samples)")
        ws = self._workspaces.get(slot)
        if ws is None or ws.numel() < need:
            self._workspaces[slot] = None
            ws = self._workspaces[slot] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws, need

    def side_streams(self, n):
        """Sub-batch streams in descending priority: the first sub-batch's kernels win every scheduling decision, so
        it finishes (and its device->host copy starts) while the second one computes, instead of both interleaving
        kernel by kernel and exposing both result copies at the end.  CUDA-graph nodes capture the priority of the
        stream they were captured on."""
        lo, hi = torch.cuda.Stream.priority_range()      # (least, greatest) = (0, -5) on current parts
        while len(self._streams) < n:
            prio = max(hi, min(lo, hi + len(self._streams)))
            self._streams.append(torch.cuda.Stream(device=self.device, priority=prio))
        return self._streams[:n]

    @staticmethod
    def _as_block(rows, max_length):
        """(B, max_length) view if the rows are consecutive full-length slices of one contiguous tensor, else None."""
        r0 = rows[0]
        if r0.dtype != torch.float32 or r0.device.type != "cpu" or not r0.is_contiguous() or r0.shape[-1] != max_length:
            return None
        p0, step = r0.data_ptr(), max_length * 4
        for i, r in enumerate(rows):
            if (r.dtype != torch.float32 or r.shape[-1] != max_length or not r.is_contiguous() or
                    r.data_ptr() != p0 + i * step or r.untyped_storage().data_ptr() != r0.untyped_storage().data_ptr()):
                return None
        return torch.as_strided(r0, (len(rows), max_length), (max_length, 1))

    def upload(self, rows, lengths, max_length, slot):
        """Host rows -> the slot's persistent (B, max_length) device batch, on the current stream.

        Rows that already sit in pinned memory are copied straight from where they are (one 2-D copy when they are
        consecutive slices of one tensor, else one copy per row); pageable rows go through the slot's pinned staging
        buffer.  Short rows are zero padded (sylber.py:107-111)."""
        B = len(rows)
        wav_dev, n_dev = self.device_input(slot, B, max_length)
        block = self._as_block(rows, max_length)
        if block is not None and block.is_pinned():
            wav_dev.copy_(block, non_blocking=True)
        elif all(r.device.type == "cpu" and r.dtype == torch.float32 and r.is_pinned() for r in rows):
            for i, r in enumerate(rows):
                k = r.shape[-1]
                wav_dev[i, :k].copy_(r.reshape(-1), non_blocking=True)
                if k < max_length:
                    wav_dev[i, k:].zero_()
        else:
            wav_dev.copy_(self.stage_input(rows, max_length, slot, block), non_blocking=True)
        n_host = torch.tensor(lengths, dtype=torch.int32)
        n_dev.copy_(n_host, non_blocking=True)
        return wav_dev, n_dev

    def _resampler(self, sample_rate):
        """Device copy of the sinc filter bank for sample_rate -> 16 kHz (cached per rate)."""
        ent = self._resamplers.get(sample_rate)
        if ent is None:
            k, width, orig_g, new_g = sinc_resample_kernel(sample_rate, 16000)
            ent = self._resamplers[sample_rate] = (torch.from_numpy(k).to(self.device).contiguous(), width, orig_g, new_g)
        return ent

    def upload_pcm16(self, rows, lengths, max_length, slot, sample_rate=16000):
        """Raw audio rows -> the slot's (B, max_length) fp32 device batch, converted, (resampled,) z-normalised and zero
        padded on the GPU (syl_prepare_pcm16 / syl_resample / syl_prepare_f32; sylber.py:83-87 + :93-118).
        Rows are int16 PCM (they travel back to back as int16: half the host->device bytes) or float32 samples in
        [-1, 1) as an audio file reader returns them (the reference's file branch).  `lengths` / `max_length` count
        16 kHz samples."""
        B = len(rows)
        is_f32 = rows[0].dtype == torch.float32
        in_len = [int(r.shape[-1]) for r in rows]
        total = sum(in_len)
        key = ("raw_f32" if is_f32 else "pcm", slot)
        buf = self._stage_in.get(key)
        if buf is None or buf[0].numel() < total or buf[2].numel() < B:
            cap = max(total, 1)
            dt = torch.float32 if is_f32 else torch.int16
            buf = self._stage_in[key] = (torch.empty(cap, dtype=dt, pin_memory=True),
                                         torch.empty(cap, dtype=dt, device=self.device),
                                         torch.empty(max(B, 64), dtype=torch.int64, device=self.device))
        host, dev, off_dev = buf
        offsets, o = [], 0
        for r, n in zip(rows, in_len):
            host[o:o + n].copy_(r)
            offsets.append(o)
            o += n
        dev[:total].copy_(host[:total], non_blocking=True)
        off_dev[:B].copy_(torch.tensor(offsets, dtype=torch.int64), non_blocking=True)
        wav_dev, n_dev = self.device_input(slot, B, max_length)
        n_dev.copy_(torch.tensor(lengths, dtype=torch.int32), non_blocking=True)
        t_in = max(in_len)
        need = int(self.lib.syl_pcm16_workspace_bytes(B, max(max_length, t_in)))
        ws = self._stage_in.get(("pcm_ws", slot))
        if ws is None or ws.numel() < need:
            ws = self._stage_in[("pcm_ws", slot)] = torch.empty(need, dtype=torch.uint8, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        prepare = self.lib.syl_prepare_f32 if is_f32 else self.lib.syl_prepare_pcm16
        pname = "syl_prepare_f32" if is_f32 else "syl_prepare_pcm16"
        if sample_rate == 16000:
            rc = prepare(_ptr(dev), _ptr(off_dev), _ptr(n_dev), B, max_length, 1, _ptr(wav_dev), _ptr(ws), need, stream)
            _lib.check(self.lib, None, rc, pname)
            return wav_dev, n_dev
        # other rates (sylber.py:84-86): -> fp32 rows, sinc resampling to 16 kHz, then (w - mean) / std - all on the device
        kern, width, orig_g, new_g = self._resampler(sample_rate)
        n_in = torch.tensor(in_len, dtype=torch.int32).to(self.device, non_blocking=True)
        raw = torch.empty((B, t_in), dtype=torch.float32, device=self.device)
        rc = prepare(_ptr(dev), _ptr(off_dev), _ptr(n_in), B, t_in, 0, _ptr(raw), _ptr(ws), need, stream)
        _lib.check(self.lib, None, rc, pname)
        res = torch.empty((B, max_length), dtype=torch.float32, device=self.device)
        rc = self.lib.syl_resample(_ptr(raw), _ptr(n_in), B, t_in, _ptr(kern), orig_g, new_g, width, _ptr(res), None, max_length, stream)
        _lib.check(self.lib, None, rc, "syl_resample")
        row_off = (torch.arange(B, dtype=torch.int64) * max_length).to(self.device, non_blocking=True)
        rc = self.lib.syl_prepare_f32(_ptr(res), _ptr(row_off), _ptr(n_dev), B, max_length, 1, _ptr(wav_dev), _ptr(ws), need, stream)
        _lib.check(self.lib, None, rc, "syl_prepare_f32")
        return wav_dev, n_dev

    def segment_states(self, states, thr_norm, thr_merge, slot=None):
        """get_segment + pooling on given (B,T,768) fp32 device states (syl_segment).  With `slot`, the outputs and the
        workspace are persistent per (slot, B, T) - overwritten by the next call of the same slot and shape."""
        B, T, _ = states.shape
        key = ("seg", slot, B, T)
        bufs = self._seg_io.get(key) if slot is not None else None
        if bufs is None:
            need = int(self.lib.syl_segment_workspace_bytes(B, T))
            bufs = (torch.empty(need, dtype=torch.uint8, device=self.device),
                    torch.empty((B, T, 2), dtype=torch.int32, device=self.device),
                    torch.empty((B,), dtype=torch.int32, device=self.device),
                    torch.empty((B, T, HIDDEN), dtype=torch.float32, device=self.device))
            if slot is not None:
                while len(self._seg_io) >= 12:
                    self._seg_io.pop(next(iter(self._seg_io)))
                self._seg_io[key] = bufs
        ws, seg, cnt, feat = bufs
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.syl_segment(_ptr(states), B, T, float(thr_norm), float(thr_merge), _ptr(seg), _ptr(cnt), _ptr(feat), T,
                                  _ptr(ws), ws.numel(), ctypes.c_void_p(stream))
        _lib.check(self.lib, None, rc, "syl_segment")
        return seg, cnt, feat

    def copy_streams(self, n):
        """One device->host copy stream per sub-batch slot: the hidden states travel while the segmentation still runs."""
        while len(self._copy_streams) < n:
            self._copy_streams.append(torch.cuda.Stream(device=self.device))
        return self._copy_streams[:n]

    def stage_input(self, rows, max_length, slot=0, block=None):
        """Zero-padded (B, max_length) fp32 batch in a persistent pinned buffer (consumed within the call)."""
        B = len(rows)
        n = B * max_length
        buf = self._stage_in.get(slot)
        if buf is None or buf.numel() < n:
            buf = self._stage_in[slot] = torch.empty(n, dtype=torch.float32, pin_memory=True)
        host = buf[:n].view(B, max_length)
        if block is not None:
            host.copy_(block)
            return host
        for i, r in enumerate(rows):
            k = r.shape[-1]
            host[i, :k].copy_(r)
            if k < max_length:
                host[i, k:].zero_()
        return host

    def device_input(self, slot, B, n):
        """Persistent device-side (B, n) fp32 batch and (B,) int32 length buffers for one slot."""
        buf = self._stage_in.get(("dev", slot))
        if buf is None or buf[0].numel() < B * n or buf[1].numel() < B:
            buf = self._stage_in[("dev", slot)] = (torch.empty(B * n, dtype=torch.float32, device=self.device),
                                                   torch.empty(max(B, 64), dtype=torch.int32, device=self.device))
        return buf[0][:B * n].view(B, n), buf[1][:B]

    def forward(self, wav, n_samples, thr_norm, thr_merge, segment=True, max_seg=None, slot=0):
        """wav (B, T_samp) fp32 cuda, n_samples (B,) int32 cuda or None.

        Returns hidden (B,T,768) and, if `segment`, (seg (B,max_seg,2) int32, seg_count (B,) int32,
        seg_feat (B,max_seg,768) fp32), all device tensors on the current stream."""
        B, t_samp = wav.shape
        T = self.num_frames(t_samp)
        ws, need = self.workspace(B, t_samp, slot)
        max_seg = (T if max_seg is None else int(max_seg)) if segment else 0
        # Output buffers are persistent per (slot, shape): stable pointers let the library replay its CUDA graph.
        # They are overwritten by the next forward of the same slot and shape - callers that keep results clone them.
        # A small LRU keeps the shapes of recurring length buckets alive.
        key = (slot, B, T, max_seg)
        io = self._io.pop(key, None)
        if io is None:
            hidden = torch.empty((B, T, HIDDEN), dtype=torch.float32, device=self.device)
            seg = cnt = feat = None
            if segment:
                seg = torch.empty((B, max_seg, 2), dtype=torch.int32, device=self.device)
                cnt = torch.empty((B,), dtype=torch.int32, device=self.device)
                feat = torch.empty((B, max_seg, HIDDEN), dtype=torch.float32, device=self.device)
            io = (hidden, seg, cnt, feat)
            while len(self._io) >= 12:
                self._io.pop(next(iter(self._io)))
        self._io[key] = io                       # most recently used last
        hidden, seg, cnt, feat = io
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.syl_forward(self.handle, _ptr(wav), _ptr(n_samples), B, t_samp, _ptr(hidden), _ptr(seg), _ptr(cnt),
                                  _ptr(feat), max_seg, float(thr_norm), float(thr_merge),
                                  _ptr(ws), need, ctypes.c_void_p(stream))
        _lib.check(self.lib, self.handle, rc, "syl_forward")
        return hidden, seg, cnt, feat

    def read_stage(self, name, shape):
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.syl_read_stage(self.handle, name.encode(), _ptr(out), out.numel(), ctypes.c_void_p(stream))
        _lib.check(self.lib, self.handle, rc, f"syl_read_stage({name})")
        return out

    def saturation_count(self):
        """How many fp16 activations of the most recent forward sit at +-65504 / are not finite (syl_saturation_scan)."""
        out = torch.zeros(1, dtype=torch.int64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.syl_saturation_scan(self.handle, _ptr(out), ctypes.c_void_p(stream))
        _lib.check(self.lib, self.handle, rc, "syl_saturation_scan")
        return int(out.item())

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
        main = torch.cuda.current_stream(eng.device)
        side = eng.side_streams(1)[0]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            hidden, _, _, _ = eng.forward(wav, n, 0.0, 0.0, segment=False, slot="speech_model")
            out = hidden.clone()
        main.wait_stream(side)
        return SimpleNamespace(last_hidden_state=out)

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
        `mode=` ("fast" default | "parity" | "strict" | "exact", see include/sylber_b200.h), `max_batch=`,
        `bucket_ratio=` (opt-in length-bucketed batching, batching.py), `thresholder=`,
        `trim_padding=True` (opt-in: padded frames of short clips are not computed - their hidden rows are zeros and
        carry no segments, the valid frames are unchanged; SYL_TRIM_PADDING in include/sylber_b200.h),
        `streams=` (sub-batches in flight, default 3: copies of one overlap kernels of the others).
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
        mode = kwargs.pop("mode", "fast")
        self.trim_padding = bool(kwargs.pop("trim_padding", False))   # opt-in: do not compute padded frames (SYL_TRIM_PADDING)
        self.max_batch = int(kwargs.pop("max_batch", 64))
        self.streams = int(kwargs.pop("streams", 3))
        self.sub_batch_sizes = kwargs.pop("sub_batch_sizes", None)
        self.bucket_ratio = kwargs.pop("bucket_ratio", None)       # e.g. 1.25: length-bucketed batching (batching.py)
        self.thresholder = kwargs.pop("thresholder", None)         # sylber_b200.Thresholder for segment(normthreshold=None)
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
        mode_bits = (_lib.MODES[mode] if isinstance(mode, str) else int(mode)) | (_lib.SYL_TRIM_PADDING if self.trim_padding else 0)
        self._engine = _Engine(state_dict, encoding_layer, device, mode_bits)
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

    @staticmethod
    def _read_files_raw(wav_file):
        """File branch with the preprocessing left to the GPU: (raw fp32 rows, common sample rate, is_batch), or None when
        the files do not share one sample rate or are not mono (the reference normalises all channels of a file jointly,
        sylber.py:86) - those go through `_prepare` on the host."""
        is_batch = isinstance(wav_file, list)
        rows, rate = [], None
        for file in (wav_file if is_batch else [wav_file]):
            w, sr = _read_audio(os.fspath(file))
            if w.dim() != 2 or w.shape[0] != 1 or (rate is not None and sr != rate):
                return None
            rate = sr
            rows.append(w[0].contiguous())
        return rows, rate, is_batch

    # ------------------------------------------------------------------------------------------
    def _sub_batches(self, n_rows):
        """(lo, hi) sub-batch bounds of one padded batch of n_rows rows (batching.sub_batch_bounds)."""
        return sub_batch_bounds(n_rows, self.streams, self.max_batch, self.sub_batch_sizes)

    def _run_jobs(self, rows, lengths, jobs, pcm=0, tables=None, after_enqueue=None, hidden_to="host"):
        """Padded batches through the engine.  rows: 1-D fp32 CPU tensors (or int16 at `pcm` Hz when pcm != 0); jobs: list of
        (row indices, max_length) - every row of a job is padded to that job's max_length (results depend on it, 8a).
        Returns per row (segments int64 (N,2) | empty, segment_features (N,768) | empty, hidden).
        `tables` = (seg (R, T, 2) int32, cnt (R,) int32) device tensors: every sub-batch also leaves its fixed-stride
        segment table there, at its rows (single-job calls only) - what segment_sharded all-gathers.  `after_enqueue(streams)`
        is called once every sub-batch has been enqueued and before the results are collected: work it enqueues behind those
        streams (the all-gather of the tables) runs while the host still waits for the last hidden states.
        `hidden_to`: "host" (the reference contract: NumPy arrays), "device" (torch CUDA tensors, no device->host copy of the
        hidden states - 49 MB per 32 x 10 s) or None (not returned).

        All sub-batches of all jobs are launched first, round-robin over the streams (a sub-batch's host->device /
        device->host copies overlap the others' kernels), and collected afterwards."""
        eng = self._engine
        work = []
        for idx, max_length in jobs:
            for lo, hi in self._sub_batches(len(idx)):
                work.append((idx[lo:hi], max_length))
        # always a side stream, even for one sub-batch: the legacy default stream cannot be graph-captured
        streams = eng.side_streams(max(1, min(self.streams, len(work))))
        main = torch.cuda.current_stream(eng.device)
        thr_n, thr_m = np.float32(self.norm_threshold), np.float32(self.merge_threshold)
        results = [None] * len(rows)

        copy_streams = eng.copy_streams(len(streams))

        def collect(item):
            st, cst, idx, hidden_h, hidden_pin, seg_pin, cnt_pin, feat, ev = item
            # the segment table travels first (it is tiny), so the number of segments is known on the host while the
            # hidden states are still in flight and the feature copy queues right behind the table: one wait, not three
            ev.synchronize()
            cnt_h = cnt_pin.numpy()
            n_max = max(int(cnt_h.max()) if len(cnt_h) else 0, 1)
            with torch.cuda.stream(st):
                feat_h, feat_pin = eng.pool.array((len(idx), n_max, HIDDEN))
                feat_pin.copy_(feat[:, :n_max], non_blocking=True)
            st.synchronize()
            cst.synchronize()
            seg_h = seg_pin.numpy()
            del hidden_pin, feat_pin
            for i, r in enumerate(idx):
                n = int(cnt_h[i])
                results[r] = (seg_h[i, :n].astype(np.int64) if n > 0 else np.array([]),
                              feat_h[i, :n] if n > 0 else np.array([]), hidden_h[i])

        pending = [None] * len(streams)                # per slot: the sub-batch whose buffers are still in use
        for k, (idx, max_length) in enumerate(work):
            slot = k % len(streams)
            st, cst = streams[slot], copy_streams[slot]
            if pending[slot] is not None:              # this slot's staging buffer, workspace and outputs are about to be reused
                collect(pending[slot])
                pending[slot] = None
            st.wait_stream(main)
            with torch.cuda.stream(st):
                chunk = [rows[i] for i in idx]
                sub_len = [lengths[i] for i in idx]
                if pcm:
                    wav_dev, n_dev = eng.upload_pcm16(chunk, sub_len, max_length, slot, pcm)
                else:
                    wav_dev, n_dev = eng.upload(chunk, sub_len, max_length, slot)
                # encoder first; its hidden states start their trip to the host on the slot's copy stream while the
                # segmentation scan (latency bound, ~0.3 ms whatever the sub-batch size) and the pooling still run
                hidden, _, _, _ = eng.forward(wav_dev, n_dev, thr_n, thr_m, segment=False, slot=slot)
                if hidden_to == "host":
                    ev_h = torch.cuda.Event()
                    ev_h.record(st)
                    cst.wait_event(ev_h)
                    with torch.cuda.stream(cst):
                        hidden_h, hidden_pin = eng.pool.array(tuple(hidden.shape))
                        hidden_pin.copy_(hidden, non_blocking=True)
                else:      # the forward's output buffer is reused by the slot's next forward: hand out a copy, or nothing
                    hidden_h, hidden_pin = (hidden.clone() if hidden_to == "device" else [None] * len(idx)), None
                seg, cnt, feat = eng.segment_states(hidden, thr_n, thr_m, slot=slot)
                if tables is not None:
                    tables[0][idx[0]:idx[-1] + 1].copy_(seg, non_blocking=True)
                    tables[1][idx[0]:idx[-1] + 1].copy_(cnt, non_blocking=True)
                cnt_pin = torch.empty(cnt.shape, dtype=torch.int32, pin_memory=True)
                cnt_pin.copy_(cnt, non_blocking=True)
                seg_pin = torch.empty(seg.shape, dtype=torch.int32, pin_memory=True)     # fixed-stride table: B x T x 2 int32
                seg_pin.copy_(seg, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
            pending[slot] = (st, cst, idx, hidden_h, hidden_pin, seg_pin, cnt_pin, feat, ev)
        if after_enqueue is not None:
            after_enqueue(streams)
        # collect in launch order: the oldest sub-batch finishes first
        n_w = len(work)
        for k in range(max(0, n_w - len(streams)), n_w):
            slot = k % len(streams)
            if pending[slot] is not None:
                collect(pending[slot])
                pending[slot] = None
        main.wait_stream(streams[0])
        return results

    @torch.no_grad()
    def __call__(self, wav_file=None, wav=None, in_second=True, pcm16=None, sample_rate=16000, pad_to=None, hidden_to="host"):
        """Same contract as the reference: a dict (single input) or list of dicts with
        `segments` (N,2), `segment_features` (N,768) float32, `hidden_states` (T_max,768) float32.

        `pcm16=` (extension): one or a list of 1-D int16 arrays / tensors of mono PCM at `sample_rate` Hz.  Equivalent
        to the reference's file branch (x / 32768, resampling to 16 kHz if needed, then (w - mean) / std,
        sylber.py:83-87) with conversion, resampling, normalisation and padding done on the GPU (syl_prepare_pcm16,
        syl_resample, syl_prepare_f32) - half the host->device bytes.

        `pad_to=` (extension): pad the batch to at least this many 16 kHz samples instead of to its own longest clip.
        An utterance's result depends on its samples and on the padded length only (conv-0 GroupNorm runs over the
        padding), so a shard of a list padded to the WHOLE list's maximum reproduces the rows of the
        single call bit for bit - this is what `sylber_b200.distributed.segment_sharded` passes.

        `hidden_to=` (extension): "host" (default, the reference contract: `hidden_states` is a NumPy array), "device"
        (`hidden_states` is a torch CUDA tensor: the 1.5 MB per 10 s clip stay on the GPU for a downstream device
        consumer) or None (`hidden_states` is None).  Segments and segment features always come back as NumPy."""
        pcm = int(sample_rate) if pcm16 is not None else 0
        if pcm:
            is_batch = isinstance(pcm16, (list, tuple))
            rows = [torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x).reshape(-1)
                    for x in (pcm16 if is_batch else [pcm16])]
            if any(r.dtype != torch.int16 for r in rows):
                raise TypeError("pcm16 expects int16 samples")
        raw_files = self._read_files_raw(wav_file) if (wav_file is not None and not pcm) else None
        if raw_files is not None:
            # the reference's file branch (sylber.py:83-87) with resampling, (w - mean) / std and padding on the GPU
            rows, pcm, is_batch = raw_files
        elif not pcm:
            batch_wavs, is_batch = self._prepare(wav_file, wav)
            rows = []
            for w in batch_wavs:
                w = torch.as_tensor(w)
                if w.dim() == 1:
                    w = w[None, :]
                rows.extend(w[i] for i in range(w.shape[0]))      # torch.cat(dim=0) at sylber.py:117: channels become rows
        lengths = [int(r.shape[-1]) for r in rows]
        if pcm and pcm != 16000:                       # lengths count 16 kHz samples from here on
            g = math.gcd(pcm, 16000)
            lengths = [resampled_length(n, pcm // g, 16000 // g) for n in lengths]
        if pad_to is not None and self.bucket_ratio:
            raise ValueError("pad_to fixes the padded length of the whole call; it cannot be combined with bucket_ratio")
        floor = int(pad_to) if pad_to is not None else 0
        if self.bucket_ratio:
            # opt-in deviation from the reference's padding semantics (batching.py): each bucket is padded to its own max
            buckets = plan_length_buckets(lengths, self.bucket_ratio, self.max_batch)
        else:
            buckets = [list(range(len(rows)))]
        results = self._run_jobs(rows, lengths, [(idx, max(floor, max(lengths[i] for i in idx))) for idx in buckets], pcm,
                                 hidden_to=hidden_to)
        outputs = [{'segments': seg * 1.0 / FRAME_RATE if in_second else seg,
                    'segment_features': feat, 'hidden_states': hid} for seg, feat, hid in results]
        return outputs if is_batch else outputs[0]

    def fp16_saturated(self):
        """Diagnostic for a new checkpoint / input domain: number of fp16 activations of the most recent forward that hit
        the saturation value (0 = every intermediate stayed inside the fp16 range, which the precision claims assume).
        Scans the workspace of the forward that ran last; not on the hot path."""
        torch.cuda.synchronize(self._engine.device)
        return self._engine.saturation_count()

    @torch.no_grad()
    def call_with_tables(self, wav, pad_to=None, after_enqueue=None, hidden_to="host"):
        """`__call__(wav=list, in_second=False, pad_to=...)` that also returns the call's fixed-stride segment table as
        DEVICE tensors, seg (B, T, 2) int32 and cnt (B,) int32 with T = frames of the padded length: the operands of the
        one collective of the sharded path (distributed.segment_sharded), gathered without a host round trip.
        `after_enqueue(streams, seg, cnt)` runs when all sub-batches are enqueued (the tables are complete once `streams`
        have drained) and before the host collects the results."""
        if self.bucket_ratio:
            raise ValueError("call_with_tables pads the whole call to one length; it cannot be combined with bucket_ratio")
        rows = []
        for w in wav:
            w = torch.as_tensor(w)
            rows.extend(w[i] for i in range(w.shape[0])) if w.dim() == 2 else rows.append(w)
        lengths = [int(r.shape[-1]) for r in rows]
        max_length = max(max(lengths), int(pad_to or 0))
        eng = self._engine
        T = eng.num_frames(max_length)
        seg = torch.empty((len(rows), T, 2), dtype=torch.int32, device=eng.device)
        cnt = torch.empty((len(rows),), dtype=torch.int32, device=eng.device)
        hook = (lambda streams: after_enqueue(streams, seg, cnt)) if after_enqueue is not None else None
        results = self._run_jobs(rows, lengths, [(list(range(len(rows))), max_length)], tables=(seg, cnt), after_enqueue=hook,
                                 hidden_to=hidden_to)
        outputs = [{'segments': sg, 'segment_features': feat, 'hidden_states': hid} for sg, feat, hid in results]
        return outputs, seg, cnt

    @torch.no_grad()
    def segment(self, input_values=None, features=None, attention_mask=None, mergethreshold=None, normthreshold=None,
                **_kw):
        """The torch-side contract of the reference's other inference caller, `Sylber.segment`
        (sylber/model/sylber.py:208-247): returns `(features, segments, avg_fts)` with features (B,T,768) on the
        device, segments a list of (N,2) int64 arrays (or the reference's empty float array) and avg_fts the per-segment
        means zero-padded to (B, max(N,1), 768).  Segmentation and pooling run on the GPU (`syl_forward` /
        `syl_segment`); thresholds default to the Segmenter's, or to `self.thresholder` when one is attached."""
        eng = self._engine
        if normthreshold is None:
            thr = getattr(self, "thresholder", None)
            normthreshold = float(thr.get_threshold()) if thr is not None else self.norm_threshold
        if mergethreshold is None:
            mergethreshold = self.merge_threshold
        thr_n, thr_m = np.float32(normthreshold), np.float32(mergethreshold)
        main = torch.cuda.current_stream(eng.device)
        side = eng.side_streams(1)[0]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if features is None:
                wav = input_values.to(device=eng.device, dtype=torch.float32).contiguous()
                n = None
                if attention_mask is not None:
                    n = attention_mask.to(eng.device).sum(-1).to(torch.int32).contiguous()
                hidden, seg, cnt, feat = eng.forward(wav, n, thr_n, thr_m, slot="segment")
                features = hidden.clone()
            else:
                features = features.to(device=eng.device, dtype=torch.float32).contiguous()
                seg, cnt, feat = eng.segment_states(features, thr_n, thr_m)
            cnt_h = cnt.cpu().numpy()
            n_max = max(int(cnt_h.max()) if len(cnt_h) else 0, 1)
            seg_h = seg[:, :n_max].cpu().numpy()
            avg = feat[:, :n_max].clone()
            keep = torch.arange(n_max, device=eng.device)[None, :] < cnt[:, None].to(torch.int64)
            avg = avg * keep[:, :, None]                       # pad_sequence(padding_value=0.0)
        main.wait_stream(side)
        segments = [seg_h[b, :int(cnt_h[b])].astype(np.int64) if cnt_h[b] > 0 else np.array([]) for b in range(len(cnt_h))]
        return features, segments, avg


class KMeansQuantizer:
    """Nearest-centroid lookup of segment features: the inference half of the reference's `KMQuantizer`
    (sylber/model/quantizer.py:86-135, `load_km_quantizer`), whose `vector_quantize_pytorch` codebook does an
    exhaustive Euclidean search.  `centroids`: path to the reference's .npy file or an array (K, 768)."""

    def __init__(self, centroids, normalize=False, device="cuda"):
        if isinstance(centroids, (str, os.PathLike)):
            centroids = np.load(centroids)
        c = torch.as_tensor(np.asarray(centroids), dtype=torch.float32)
        if c.dim() != 2 or c.shape[1] != HIDDEN:
            raise ValueError(f"centroids must be (K, {HIDDEN}), got {tuple(c.shape)}")
        if not torch.cuda.is_available():
            raise RuntimeError("sylber_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load_library()
        self.device = torch.device(device if "cuda" in str(device) else "cuda")
        self.centroids = c.to(self.device).contiguous()
        self.normalize = bool(normalize)

    @torch.no_grad()
    def get_indices(self, token, return_distance=False):
        """token (..., 768) array or tensor -> int64 indices of shape token.shape[:-1] + (1,): the reference returns
        `indices[0]` of a GroupedResidualVQ with one quantizer, i.e. a trailing quantizer axis of length 1
        (quantizer.py:95-106), and its callers index it."""
        t = torch.as_tensor(token)
        lead = tuple(t.shape[:-1])
        x = t.reshape(-1, HIDDEN).to(device=self.device, dtype=torch.float32).contiguous()
        n = x.shape[0]
        idx = torch.empty((n,), dtype=torch.int32, device=self.device)
        dist = torch.empty((n,), dtype=torch.float32, device=self.device) if return_distance else None
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.syl_kmeans_assign(_ptr(x), n, _ptr(self.centroids), self.centroids.shape[0], int(self.normalize),
                                            _ptr(idx), _ptr(dist), ctypes.c_void_p(stream))
        _lib.check(self.lib, None, rc, "syl_kmeans_assign")
        out = idx.to(torch.int64).reshape(lead + (1,))
        return (out, dist.reshape(lead)) if return_distance else out

    def decode(self, indices):
        """quantizer.py:123-129: centroids of `indices[..., :1]` (negative indices clipped to 0) -> (..., 768).
        Accepts the (..., 1) layout get_indices returns as well as plain (...) index arrays."""
        idx = torch.as_tensor(indices, device=self.device)
        if idx.dim() >= 1 and idx.shape[-1] == 1:
            idx = idx[..., 0]
        return self.centroids[idx.clip(0).long()]
