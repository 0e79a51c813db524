#!/usr/bin/env python
"""Benchmark of the Segmenter forward path (BASELINE.json: audio frames/s on 10 s, 16 kHz clips).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--layers 9] [--mode parity]
                    [--workload 10s|60s|mixed]

One "step" = one pass of the hot path (conv front end -> 9-layer encoder -> segmentation -> segment pooling)
over one batch of synthetic audio.  Workloads, per GPU (weak scaling, the rank's utterances are its own):
  10s    BASELINE.json configs[1] / [2]: batch 32 x 10 s - the configuration the metric is quoted on (default)
  60s    configs[3]: batch 8 x 60 s, T = 2999 - the attention-roofline case
  mixed  configs[4]: 64 clips of 2-30 s (lengths torch.randint(32000, 480001, seed 2)), zero padded to
         the GLOBAL maximum over all ranks' clips - the T_max one process calling the reference on the whole list
         would use (sylber.py:93-118; results depend on it).  frames/s counts VALID frames.
Audio is N(0,1) (the distribution after the reference's (w-mean)/std); hubert-base architecture with the reference's 9
encoder layers, synthetic weights with speech-like segment occupancy (weights.SPEECH_LIKE_BIAS_NORM; no checkpoint
exists offline; SYLBER_CKPT=<path> uses a real one).

JSON line (rank 0): value = frames/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = the same through the public call (Segmenter.__call__, or segment_sharded with its all-gather at N > 1) from
pinned host tensors to NumPy results; roofline = the dominant kernel (gemm3_tc_kernel) TIME-WEIGHTED over all its
launches against MEASURED_PEAKS.json, with the best launch, the attention+MLP path (LayerNorms included) and the
whole step as sub-fields; stages = per-stage device time and achieved rates; cpu_baseline = the reference's own CPU
implementation (oracle/_ref, the unmodified sylber.Segmenter) on this box's host cores, which also yields
config.segment_agreement (device segments vs the reference's on the same clips).
`--impl reference` times that CPU path alone on the same clips.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "audio frames/sec (16 kHz, 10 s clips)"
UNIT = "frames/s"
THR_NORM, THR_MERGE = 2.6, 0.8
CONV_K = (10, 3, 3, 3, 3, 2, 2)
CONV_S = (5, 2, 2, 2, 2, 2, 2)
MODE_ERR = {"parity": "conv4-6 and the feature projection run split hi/lo f16 = 3 passes; 4.1e-4 rel vs fp32",
            "strict": "conv1-6, projection, pos-conv split; 3.0e-4", "exact": "every GEMM split; 2.5e-5",
            "fast": "no split; 5.1e-4"}
SPLIT_CONV = {"fast": (), "parity": (4, 5, 6), "strict": (1, 2, 3, 4, 5, 6), "exact": (1, 2, 3, 4, 5, 6)}


def conv_lengths(n):
    out = []
    for k, s in zip(CONV_K, CONV_S):
        n = (n - k) // s + 1
        out.append(n)
    return out


def stage_flops(n_samples, layers):
    """Algorithmic FLOPs (2*MAC) per utterance and stage."""
    L = conv_lengths(n_samples)
    T = L[6]
    return {
        "conv0_gn_gelu": 2 * 512 * 10 * L[0],
        "conv1_gemm": 2 * 512 * 512 * 3 * L[1],
        "conv2_6_gemm": 2 * 512 * 512 * (3 * (L[2] + L[3] + L[4]) + 2 * (L[5] + L[6])),
        "feature_proj_gemm": 2 * T * 512 * 768,
        "pos_conv_gemm": 2 * T * 768 * 48 * 128,
        "qkv_gemm": layers * 2 * T * 768 * 2304,
        "attention": layers * 4 * T * T * 768,
        "out_proj_gemm": layers * 2 * T * 768 * 768,
        "ffn1_gemm": layers * 2 * T * 768 * 3072,
        "ffn2_gemm": layers * 2 * T * 768 * 3072,
    }, T, L


def gemm_algorithmic_bytes(B, L, T, layers, mode):
    """Algorithmic DRAM bytes of all gemm3_tc_kernel launches of one step (operands read once, outputs written once):
    fp16 operands are 2 bytes, 4 where the site runs split (hi + lo), fp32 outputs 4 bytes."""
    sc = set(SPLIT_CONV[mode])
    s_proj = 2 if mode in ("parity", "strict", "exact") else 1
    s_pos = 2 if mode in ("strict", "exact") else 1
    s_enc = 2 if mode == "exact" else 1
    M = B * T
    total = 0
    for i, k in zip(range(1, 7), (3, 3, 3, 3, 2, 2)):
        s_in = 2 if i in sc else 1
        total += B * L[i - 1] * 512 * 2 * s_in + 512 * 512 * k * 2 * s_in
        total += B * L[i] * 512 * (4 if i == 6 else 2 * (2 if (i + 1) in sc else 1))
    total += M * 512 * 2 * s_proj + 768 * 512 * 2 * s_proj + M * 768 * (4 + 2 * s_pos)
    per_layer = (M * 768 * 2 * s_enc + 2304 * 768 * 2 * s_enc + M * 2304 * 2          # QKV
                 + M * 768 * 2 * s_enc + 768 * 768 * 2 * s_enc + M * 768 * 4           # out-projection
                 + M * 768 * 2 * s_enc + 3072 * 768 * 2 * s_enc + M * 3072 * 2 * s_enc  # FFN1
                 + M * 3072 * 2 * s_enc + 768 * 3072 * 2 * s_enc + M * 768 * 4)         # FFN2
    return total + layers * per_layer


GEMM_STAGES = ("conv1_gemm", "conv2_6_gemm", "feature_proj_gemm", "qkv_gemm", "out_proj_gemm", "ffn1_gemm", "ffn2_gemm")
ENC_STAGES = ("qkv_gemm", "attention", "out_proj_gemm", "ffn1_gemm", "ffn2_gemm")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed regions."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def make_workload(name, world, rank, pin=True):
    """(list of this rank's (1, n) fp32 clips, padded length of the call, description).  Deterministic in (name, world)."""
    pin = pin and torch.cuda.is_available()
    if name in ("10s", "60s"):
        n, B = (160000, 32) if name == "10s" else (960000, 8)
        g = torch.Generator().manual_seed(1)
        if world * B <= 256:
            wav = torch.randn(B * world, n, generator=g)[rank * B:(rank + 1) * B].contiguous()
        else:
            wav = torch.randn(B, n, generator=torch.Generator().manual_seed(1 + rank))
        wav = wav.pin_memory() if pin else wav       # e2e inputs start in pinned host memory
        return [wav[i:i + 1] for i in range(B)], n, f"batch={B} synthetic {n // 16000} s 16 kHz wav per GPU"
    B = 64
    lens = torch.randint(32000, 480001, (B * world,), generator=torch.Generator().manual_seed(2)).tolist()
    g = torch.Generator().manual_seed(1 + rank)
    mine = lens[rank * B:(rank + 1) * B]
    clips = [torch.randn(1, n, generator=g) for n in mine]
    clips = [c.pin_memory() for c in clips] if pin else clips
    return clips, max(lens), (f"{B} clips of 2-30 s per GPU (lengths torch.randint(32000, 480001, ({B * world},), seed 2)), every "
                              f"clip zero padded to the GLOBAL maximum over all ranks ({max(lens)} samples) = the reference's "
                              f"padding semantics for one call on the whole list")


def load_weights(layers):
    from sylber_b200.weights import syllabic_test_state_dict, normalize_state_dict, SPEECH_LIKE_BIAS_NORM
    ckpt = os.environ.get("SYLBER_CKPT")
    if ckpt:
        return normalize_state_dict(torch.load(ckpt, map_location="cpu")), f"checkpoint {ckpt}"
    return syllabic_test_state_dict(layers, 0, bias_norm=SPEECH_LIKE_BIAS_NORM), "synthetic"


# --------------------------------------------------------------------------------------------------
# CPU path: the unmodified reference from oracle/_ref when it is installed, else the oracle port
# --------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, sd, layers):
        from oracle import make_ref
        torch.set_num_threads(os.cpu_count())
        self.sd, self.layers = sd, layers
        self.kind = "reference" if make_ref.available() else "port"
        if self.kind == "reference":
            self.seg = make_ref.reference_segmenter(sd, layers)
            _, self.get_segment = make_ref.load_reference()
        else:
            from oracle import segment_ref
            self.get_segment = segment_ref.get_segment

    def full(self, clips):
        """What the user of the reference runs: Segmenter.__call__(wav=list) -> list of dicts."""
        if self.kind == "reference":
            return self.seg(wav=clips, in_second=False)
        from oracle.hubert_ref import hubert_forward
        from oracle import segment_ref
        lens = [c.shape[1] for c in clips]
        batch = torch.zeros(len(clips), max(lens))
        for i, c in enumerate(clips):
            batch[i, :lens[i]] = c[0]
        hidden = hubert_forward(self.sd, batch, lens, self.layers).numpy()
        return [segment_ref.package(st, segment_ref.get_segment(st, THR_NORM, THR_MERGE), in_second=False) for st in hidden]

    def split(self, clips):
        """(model-only ms, segmentation-only ms) of one pass: the HubertModel forward and the get_segment loop apart."""
        lens = [c.shape[1] for c in clips]
        batch = torch.zeros(len(clips), max(lens))
        mask = torch.zeros(len(clips), max(lens), dtype=torch.long)
        for i, c in enumerate(clips):
            batch[i, :lens[i]] = c[0]
            mask[i, :lens[i]] = 1
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                hidden = self.seg.speech_model(batch, attention_mask=mask).last_hidden_state.numpy()
            else:
                from oracle.hubert_ref import hubert_forward
                hidden = hubert_forward(self.sd, batch, lens, self.layers).numpy()
        t1 = time.perf_counter()
        for st in hidden:
            self.get_segment(st, THR_NORM, THR_MERGE)
        t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3

    def describe(self):
        return ("unmodified sylber.Segmenter from oracle/_ref (pip-installed copy of the reference; HubertModel of the installed "
                "transformers, NumPy get_segment), device='cpu'" if self.kind == "reference" else
                "oracle/ port (torch CPU fp32 ops + NumPy get_segment restatement)")


def cpu_sample(name, clips):
    """Bounded sample of the rank's clips the CPU path is timed on (the whole batch where that is a few seconds)."""
    if name == "10s":
        return clips, f"all {len(clips)} clips of the step"
    if name == "60s":
        return clips[:4], "4 of the step's 8 clips"
    return clips[:16], "the first 16 of the step's 64 clips, padded to their own maximum (reference list call)"


def valid_frames(clips):
    return sum(conv_lengths(c.shape[1])[6] for c in clips)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd, wsrc = load_weights(args.layers)
    clips, pad_to, desc = make_workload(args.workload, 1, 0, pin=False)
    sample, sample_desc = cpu_sample(args.workload, clips)
    cpu = CpuPath(sd, args.layers)
    frames = valid_frames(sample)
    for _ in range(max(args.warmup, 0)):
        cpu.full(sample)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cpu.full(sample)
        times.append(time.perf_counter() - t0)
    ms = sum(times) / len(times) * 1e3
    fps = frames / (ms / 1e3)
    model_ms, seg_ms = cpu.split(sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{desc}, sylber_base ({args.layers}L/768d)", "weights": wsrc,
                   "note": f"CPU arm: {cpu.describe()}; each step = Segmenter.__call__ on {sample_desc}"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu.kind,
                         "sample": f"{sample_desc} ({frames} valid frames per step), {args.steps} steps, torch {torch.__version__} fp32, "
                                   f"os.cpu_count()={os.cpu_count()}",
                         "split_ms": {"model_only": round(model_ms, 1), "segmentation_only": round(seg_ms, 1), "full": round(ms, 1)}},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# the PyTorch-eager library bar on the same GPU: transformers.HubertModel, model forward only
# --------------------------------------------------------------------------------------------------
def library_baseline(sd, layers, wav_dev, T, reps=3):
    try:
        from transformers import HubertConfig, HubertModel
        model = HubertModel(HubertConfig(num_hidden_layers=layers))
        model.load_state_dict(sd, strict=False)
        model = model.eval().to(wav_dev.device)
        out = {}
        for name, ctx, tf32 in (("fp32", None, False), ("tf32", None, True), ("bf16_autocast", torch.bfloat16, False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad(), torch.autocast("cuda", dtype=ctx, enabled=ctx is not None):
                model(wav_dev)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    model(wav_dev)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"ms_per_forward": round(ms, 2), "frames_per_s": round(wav_dev.shape[0] * T / (ms / 1e3))}
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        del model
        torch.cuda.empty_cache()
        out["what"] = ("transformers.HubertModel eager on this GPU (cuDNN convs, cuBLAS linears, SDPA), model forward only - no "
                       "segmentation / pooling; fp32 is what the reference runs with device='cuda'")
        return out
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from sylber_b200 import Segmenter, segment_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity") and not args.no_affinity:
        # one process per GPU on one host: give every rank its own share of the host cores, so that the ranks' launch /
        # packaging threads do not migrate over each other (measured: e2e at 4 ranks, profiles/r03_multi_gpu.md)
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 1:
            mine_cores = cores[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine_cores)
            torch.set_num_threads(max(1, min(per, 4)))
            affinity = f"{per} of {len(cores)} host cores per rank"
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set; keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    layers = args.layers
    sd, wsrc = load_weights(layers)
    clips, pad_to, desc = make_workload(args.workload, world, rank)
    B = len(clips)
    max_batch = 32
    seg = Segmenter(model_ckpt=None, state_dict=sd, encoding_layer=layers, device=f"cuda:{local}", mode=args.mode,
                    max_batch=max_batch, trim_padding=args.trim, **({"streams": args.streams} if args.streams else {}))
    eng = seg._engine
    lens = [c.shape[1] for c in clips]
    flops, T, L = stage_flops(pad_to, layers)           # executed work: every clip is padded to pad_to
    if args.trim:
        # trimmed mode computes only an utterance's own frames (keys are masked to them already): the executed work per
        # clip is that of its own length, averaged here so that `flops[stage] * B` stays the step's total
        per_clip = [stage_flops(n, layers)[0] for n in lens]
        flops = {k: sum(f[k] for f in per_clip) / len(per_clip) for k in flops}
    frames_valid_local = valid_frames(clips)
    thr_n, thr_m = np.float32(THR_NORM), np.float32(THR_MERGE)

    # device-resident inputs: sub-batches of <= max_batch rows, each with its own workspace / output slot
    subs = []
    for lo in range(0, B, max_batch):
        hi = min(lo + max_batch, B)
        wav_dev = torch.zeros(hi - lo, pad_to, device=dev)
        for i in range(lo, hi):
            wav_dev[i - lo, :lens[i]] = clips[i][0].to(dev)
        n_dev = torch.tensor(lens[lo:hi], dtype=torch.int32, device=dev)
        subs.append((wav_dev, n_dev))
    per_rank = B
    # the one exchange of the path: (count, fixed-stride segment table) of every utterance as ONE (B, 1 + T, 2) int32 block
    gathered = torch.empty((world * per_rank, 1 + T, 2), dtype=torch.int32, device=dev) if world > 1 else None

    run_stream = torch.cuda.Stream(device=dev)      # a real stream: the library replays its CUDA graph on it
    torch.cuda.set_stream(run_stream)

    def step():
        outs = [eng.forward(w, n, thr_n, thr_m, slot=("bench", k)) for k, (w, n) in enumerate(subs)]
        if world > 1:  # the fixed-stride segment table, count in row 0: one all-gather over NCCL / NVSwitch
            block = torch.cat([torch.cat([o[2].view(-1, 1, 1).expand(-1, 1, 2), o[1]], dim=1) for o in outs])
            dist.all_gather_into_tensor(gathered, block)
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    frames_valid = sum_over_ranks(frames_valid_local)

    # ---------------- device-resident timing (value) ----------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    # per-stage device time: the same K steps again with the library's stage events enabled (eager launches instead
    # of the CUDA-graph replay used above, because events cannot be read back from inside a graph)
    eng.profile(True)
    eng.profile_read()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for _ in range(args.steps):
        step()
    p1.record()
    barrier()
    ms_prof_total = p0.elapsed_time(p1)
    prof = eng.profile_read()
    eng.profile(False)
    value = frames_valid * args.steps / (ms_total / 1e3)
    seg_counts = torch.cat([o[2] for o in out]).cpu().numpy()

    # The attention + MLP path inside the CUDA-graph replay (the launch mode `value` is measured in): the forward without
    # segmentation at all encoder layers minus the same forward at 0 layers = what the layers add to the step.  The stage
    # timers above bracket eager launches with events, which adds ~10 us of idle per kernel to every stage they report.
    def graph_forward_ms(active):
        eng.set_active_layers(active)
        f = lambda: [eng.forward(w, n, thr_n, thr_m, segment=False, slot=("bench_enc", k)) for k, (w, n) in enumerate(subs)]
        for _ in range(3):
            f()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(args.steps):
            f()
        g1.record()
        torch.cuda.synchronize()
        return g0.elapsed_time(g1) / args.steps
    enc_graph_ms = None
    if rank == 0:
        t_all = graph_forward_ms(layers)
        t_none = graph_forward_ms(0)
        enc_graph_ms = t_all - t_none
    eng.set_active_layers(layers)

    # ---------------- end-to-end through the public call, from pinned host tensors to NumPy results ----------------
    def e2e_call():
        if world > 1:   # product-level sharded call: local forward + all-gather of the segment table on every step
            return segment_sharded(seg, wav=clips, in_second=True, local_input=True, pad_to=pad_to, per_rank=B)
        return seg(wav=clips, in_second=True, pad_to=pad_to)

    for _ in range(3):
        res = e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_call()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_local_s = None
    if world > 1:
        # the same call without the exchange (every rank only its own clips): separates the collective + straggler cost
        # from what N processes sharing one host cost each other
        for _ in range(2):
            seg(wav=clips, in_second=True, pad_to=pad_to)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            seg(wav=clips, in_second=True, pad_to=pad_to)
        torch.cuda.synchronize()
        e2e_local_s = max_over_ranks(time.perf_counter() - t0)
    e2e_dev_s = None
    if world > 1:
        # the sharded call with the hidden states left on the GPUs (hidden_to="device"): what a multi-GPU pipeline that
        # consumes segments / features pays - the host link is shared by all ranks of the box (profiles/r03_multi_gpu.md)
        f = lambda: segment_sharded(seg, wav=clips, in_second=True, local_input=True, pad_to=pad_to, per_rank=B, hidden_to="device")
        for _ in range(2):
            f()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            f()
        torch.cuda.synchronize()
        e2e_dev_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = frames_valid * args.steps / e2e_s
    mine = res[rank * B:(rank + 1) * B] if world > 1 else res
    h2d = sum(lens) * 4 + B * 4
    d2h = B * T * 768 * 4 + B * 4 + B * T * 2 * 4 + sum(int(np.asarray(r["segment_features"]).size) * 4 for r in mine)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline bookkeeping ----------------
    peaks = load_peaks()
    stages = {}
    for name, (ms, cnt) in prof.items():
        if cnt == 0:
            continue
        ms_step = ms / args.steps
        entry = {"ms_per_step": round(ms_step, 4), "share": round(ms / ms_prof_total, 4), "launch_regions": cnt // args.steps}
        if name in flops and name != "conv0_gn_gelu":
            tf = flops[name] * B / (ms_step * 1e-3) / 1e12
            entry.update({"bound": "tensor", "achieved_tflops": round(tf, 1), "frac": round(tf / peaks["tf_sustained"], 4)})
        stages[name] = entry
    M = B * T
    if "conv0_gn_gelu" in stages:
        n_out = 2 if args.mode in ("strict", "exact") else 1     # fp16 hi (+ lo only when conv1 runs split)
        by = B * (2 * pad_to * 4 + L[0] * 512 * 2 * n_out)       # wav read twice (stats + apply), fp16 activation written
        gbs = by / (stages["conv0_gn_gelu"]["ms_per_step"] * 1e-3) / 1e9
        stages["conv0_gn_gelu"].update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 4)})
    if "layernorm" in stages:    # LN(512): fp32 in, fp16 hi (+ lo) out; LN(768) after the positional conv: two fp32 in, fp16 pair out
        by = M * 512 * (4 + (4 if args.mode != "fast" else 2)) + M * 768 * (4 + 4 + 4)
        gbs = by / (stages["layernorm"]["ms_per_step"] * 1e-3) / 1e9
        stages["layernorm"].update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 4)})
    if "layernorm_encoder" in stages:   # fp32 GEMM output + residual pair in, pair out; the last one also writes fp32 hidden
        by = 2 * layers * M * 768 * (4 + 4 + 4) + M * 768 * 4
        gbs = by / (stages["layernorm_encoder"]["ms_per_step"] * 1e-3) / 1e9
        stages["layernorm_encoder"].update({"bound": "hbm", "achieved_gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 4)})
    if "attention" in stages:
        by = layers * B * 4 * T * 768 * 2
        stages["attention"]["hbm_gbs"] = round(by / (stages["attention"]["ms_per_step"] * 1e-3) / 1e9, 1)
        stages["attention"]["hbm_frac"] = round(stages["attention"]["hbm_gbs"] / peaks["hbm_gbs"], 4)

    def tf_of(names, extra_ms=0.0):
        ms = sum(stages[s]["ms_per_step"] for s in names if s in stages) + extra_ms
        fl = sum(flops[s] for s in names if s in flops) * B
        return (fl / (ms * 1e-3) / 1e12 if ms else 0.0), ms, fl

    # Dominant kernel = gemm3_tc_kernel: conv1-6 (implicit GEMM), feature projection, QKV, out-projection, FFN1, FFN2 -
    # 7 + 4 * layers launches per step.  `achieved` is TIME-WEIGHTED over all of them (sum of algorithmic FLOPs / sum of
    # their device time); the best single launch (conv1, which has its own stage timer) is a sub-field.
    dom_tf, dom_ms, dom_fl = tf_of(GEMM_STAGES)
    n_launch = 7 + 4 * layers
    traffic = alg_bytes = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tsrc = None
    if os.path.exists(tpath) and args.workload == "10s" and args.mode == "fast" and layers == 9:
        tj = json.load(open(tpath))
        ent = tj.get("gemm3_tc_kernel.all_launches") or {}
        traffic, tsrc = ent.get("dram_bytes_per_launch_mean"), ent.get("source")
    alg_bytes = gemm_algorithmic_bytes(B, L, T, layers, args.mode) / n_launch
    enc_tf, enc_ms, _ = tf_of(ENC_STAGES)
    enc_ln_ms = stages.get("layernorm_encoder", {}).get("ms_per_step", 0.0)
    encln_tf, encln_ms, _ = tf_of(ENC_STAGES, enc_ln_ms)
    step_ms = ms_total / args.steps
    all_fl = sum(flops.values()) * B
    step_tf = all_fl / (step_ms * 1e-3) / 1e12
    per_layer = [2 * 512 * 512 * k * L[i] for i, k in zip(range(1, 7), (3, 3, 3, 3, 2, 2))]
    conv_ms = stages["conv1_gemm"]["ms_per_step"] + stages["conv2_6_gemm"]["ms_per_step"]
    conv_eq = sum(f * (3 if i in SPLIT_CONV[args.mode] else 1) for i, f in zip(range(1, 7), per_layer)) * B / (conv_ms * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "kernel": "gemm3_tc_kernel",
        "launch": f"time-weighted over all {n_launch} launches per step (conv1-6, feature projection, QKV, out-proj, FFN1, FFN2)",
        "achieved": round(dom_tf, 1), "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": round(dom_tf / peaks["tf_sustained"], 4),
        "traffic": traffic,
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); fp16 and bf16 share the tensor rate",
        "algorithmic_flops_per_launch": dom_fl / n_launch, "launches_per_step": n_launch,
        "kernel_ms_per_step": round(dom_ms, 4), "kernel_share_of_step": round(dom_ms / (ms_prof_total / args.steps), 4),
        "algorithmic_bytes_per_launch": alg_bytes, "traffic_source": tsrc,
        "best_launch": {"launch": f"conv1 implicit GEMM, M={B}x{L[1]} N=512 K=1536", "achieved": stages["conv1_gemm"]["achieved_tflops"],
                        "frac": stages["conv1_gemm"]["frac"], "ms": stages["conv1_gemm"]["ms_per_step"]},
        "per_stage_frac": {s: stages[s]["frac"] for s in GEMM_STAGES if s in stages},
        "conv_stack_tensor_work": {"achieved_incl_split_passes": round(conv_eq, 1), "unit": "TFLOP/s issued to the tensor cores",
                                   "frac": round(conv_eq / peaks["tf_sustained"], 4), "ms_per_step": round(conv_ms, 4)},
        "attn_mlp_path": {"achieved": round(encln_tf, 1), "frac": round(encln_tf / peaks["tf_sustained"], 4), "unit": "TFLOP/s",
                          "ms_per_step": round(encln_ms, 4),
                          "what": "QKV + attention + out-proj + FFN1 + FFN2 + the two LayerNorms of every encoder layer",
                          "without_layernorms": {"achieved": round(enc_tf, 1), "frac": round(enc_tf / peaks["tf_sustained"], 4),
                                                 "ms_per_step": round(enc_ms, 4)},
                          "timing": "per-stage CUDA events around EAGER launches (a profiled pass after the timed region)"},
        "attn_mlp_path_in_graph": {
            "achieved": round(sum(flops[s_] for s_ in ENC_STAGES if s_ in flops) * B / (enc_graph_ms * 1e-3) / 1e12, 1),
            "frac": round(sum(flops[s_] for s_ in ENC_STAGES if s_ in flops) * B / (enc_graph_ms * 1e-3) / 1e12 / peaks["tf_sustained"], 4),
            "unit": "TFLOP/s", "ms_per_step": round(enc_graph_ms, 4),
            "what": "the same path (LayerNorms included) as the CUDA-graph replay runs it: device time of the forward at all "
                    "encoder layers minus the forward at 0 layers, no segmentation in either - the launch mode `value` is "
                    "measured in"} if enc_graph_ms else None,
        "step": {"achieved": round(step_tf, 1), "frac": round(step_tf / peaks["tf_sustained"], 4), "unit": "TFLOP/s",
                 "what": "all algorithmic FLOPs of the step / device-resident step time (segmentation, LayerNorms, conv0 included in the time)"},
        "attention_hbm": {"achieved": stages.get("attention", {}).get("hbm_gbs"), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": stages.get("attention", {}).get("hbm_frac")},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"f16 tensor-core operands, f32 accumulate/LayerNorm/softmax/residual; mode={args.mode} ({MODE_ERR[args.mode]})",
        "data": "synthetic",
        "config": {"workload": f"{desc}, sylber_base ({layers}L/768d), 1xB200 per rank", "weights": wsrc,
                   "frames_padded_per_clip": T, "clips_per_gpu": B, "valid_frames_per_step": int(frames_valid),
                   "padded_frames_per_step": int(world * B * T), "t_max_definition": "global maximum over all ranks' clips",
                   "mode": args.mode, "trim_padding": bool(args.trim), "parallelism": f"dp{world} by utterance", "host_affinity": affinity,
                   "l2": "per-step working set (~4 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "segments_per_clip_mean": float(seg_counts.mean()),
                   "e2e_input": f"list of {B} (1, n) fp32 views of pinned host memory"
                                + ("; e2e = segment_sharded(local_input=True): forward + all-gather of counts and segment table" if world > 1 else "")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3,
                **({"ms_per_step_without_the_all_gather": e2e_local_s / args.steps * 1e3} if e2e_local_s else {}),
                **({"hidden_states_left_on_device": {"value": frames_valid * args.steps / e2e_dev_s, "ms_per_step": e2e_dev_s / args.steps * 1e3,
                                                     "d2h_bytes_per_step": d2h - B * T * 768 * 4}} if e2e_dev_s else {})},
        "gpu_launches": eng.launch_count(True) * len(subs) * args.steps,
        "clocks": clocks,
        "roofline": roofline,
        "stages": stages,
        "stages_note": f"stage times from {args.steps} extra steps with per-stage CUDA events (eager launches, "
                       f"{ms_prof_total / args.steps:.3f} ms/step); the timed region above replays the same launches as a CUDA graph",
    }
    if world == 1 and not args.no_cpu:
        torch.cuda.synchronize()
        from oracle import agreement as A
        cpu = CpuPath(sd, layers)
        sample, sample_desc = cpu_sample(args.workload, clips)
        n_s = len(sample)
        ours = seg(wav=sample, in_second=False)            # the device result on exactly the clips the CPU path sees
        cpu.full(sample)
        times = []
        for _ in range(2):
            t0 = time.perf_counter()
            ref = cpu.full(sample)
            times.append(time.perf_counter() - t0)
        ms = sum(times) / len(times) * 1e3
        model_ms, seg_ms = cpu.split(sample)
        line["cpu_baseline"] = {"value": valid_frames(sample) / (ms / 1e3), "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu.kind,
                                "sample": f"{sample_desc}: 1 warm-up + 2 timed passes of {cpu.describe()}, os.cpu_count()={os.cpu_count()}",
                                "ms_per_step": ms,
                                "split_ms": {"model_only": round(model_ms, 1), "segmentation_only": round(seg_ms, 1), "full": round(ms, 1)}}
        recs = [A.compare_utterance(ref[i]["hidden_states"], ours[i]["hidden_states"], ours[i]["segments"], THR_NORM, THR_MERGE)
                for i in range(n_s)]
        summ = A.summarize(recs)
        line["config"]["segment_agreement"] = {
            "clips_with_identical_segments": summ["agree"], "clips": n_s, "mode": args.mode,
            "all_differences_explained_by_margin_below_state_error": summ["all_flips_explained"],
            "hidden_rel_err_max": summ["rel_max"], "min_margin": summ["min_margin"],
            "flips": [{k: f[k] for k in ("utterance", "kind", "frame", "margin", "delta")} for f in summ["flips"]],
            "note": "device segments vs the CPU reference's on the same clips; get_segment thresholds fp32 norms / cosines, so a clip "
                    "differs exactly when one of its decisions sits closer to the threshold than the state error (oracle/agreement.py)"}
    if world == 1 and args.library_baseline:
        line["gpu_library_baseline"] = library_baseline(sd, layers, subs[0][0], T)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=9)
    ap.add_argument("--mode", default="fast", choices=["parity", "strict", "fast", "exact"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / segment_agreement leg")
    ap.add_argument("--library-baseline", action="store_true", help="also time transformers.HubertModel eager on the GPU")
    ap.add_argument("--no-affinity", action="store_true", help="N > 1: do not pin the ranks to disjoint host-core sets")
    ap.add_argument("--trim", action="store_true", help="trimmed mode: padded frames are not computed (opt-in deviation)")
    ap.add_argument("--streams", type=int, default=0, help="sub-batches in flight in the e2e leg (0 = Segmenter default)")
    ap.add_argument("--workload", default="10s", choices=["10s", "60s", "mixed"],
                    help="10s = BASELINE configs[1]/[2] (batch 32 x 10 s per GPU, the metric's configuration); "
                         "60s = configs[3] (batch 8 x 60 s, T = 2999); mixed = configs[4] (64 clips of 2-30 s per GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
