"""BASELINE configs[4]: mixed 2-30 s clips (segment-pool + padding stress).  End-to-end frames/s on VALID frames
through Segmenter.__call__, with the reference's padding semantics (every clip padded to the batch maximum) and with
opt-in length bucketing (each bucket padded to its own maximum; results = one reference call per bucket).
    python tools/bench_mixed.py [B=64] [steps=5]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter, plan_length_buckets
from sylber_b200.batching import padded_work
from sylber_b200.weights import syllabic_test_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lens = torch.randint(32000, 480001, (B,), generator=torch.Generator().manual_seed(2)).tolist()
g = torch.Generator().manual_seed(1)
wavs = [torch.randn(1, n, generator=g).pin_memory() for n in lens]
sd = syllabic_test_state_dict(9, 0)
out = {"workload": f"{B} clips of 2-30 s (seeded), 9 layers, parity mode", "valid_seconds": sum(lens) / 16000.0}
for name, kw in (("padded_reference_semantics", {}), ("bucketed_1.25", {"bucket_ratio": 1.25})):
    seg = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", max_batch=32, **kw)
    frames = sum(seg._engine.num_frames(n) for n in lens)
    for _ in range(2):
        res = seg(wav=wavs, in_second=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        res = seg(wav=wavs, in_second=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
    buckets = plan_length_buckets(lens, 1.25, 32) if kw else None
    out[name] = {"ms_per_call": round(dt * 1e3, 2), "valid_frames_per_s": round(frames / dt),
                 "padded_samples_over_valid": round(padded_work(lens, buckets) / sum(lens), 3),
                 "segments_total": int(sum(len(r["segments"]) for r in res))}
    del seg
    torch.cuda.empty_cache()
print(json.dumps(out))
