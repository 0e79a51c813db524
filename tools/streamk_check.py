"""A/B check of the stream-K GEMM schedule (sylber_b200/csrc/gemm3_tc.cuh, SYL_STREAMK=1) on one GPU.

    python tools/streamk_check.py            # parent: runs the two children below under a timeout and compares
    python tools/streamk_check.py child OUT  # child: one process, schedule chosen by SYL_STREAMK in its environment

The children run batch 32 x 10 s (192 tiles for the N = 768 GEMMs, the case the schedule is for) three times - eager
launch, then CUDA-graph replays - and save hidden states and segments.  The parent requires: the stream-K runs are
bit-identical to each other (the partial sums are added in a fixed order), hidden states within 1e-5 relative of the
data-parallel schedule, and reports how many utterances changed a segment boundary (a different summation order can
move a threshold decision; the count should be zero or close to it) plus the per-stage device times of both."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(out):
    import numpy as np
    import torch
    from sylber_b200 import Segmenter
    from sylber_b200.weights import syllabic_test_state_dict
    sd = syllabic_test_state_dict(9, seed=0)
    seg = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", streams=1)
    g = torch.Generator().manual_seed(1)
    wav = torch.randn(32, 160000, generator=g)
    rows = [wav[i:i + 1] for i in range(32)]
    runs = []
    for _ in range(3):
        res = seg(wav=rows, in_second=False)
        runs.append((np.stack([r["hidden_states"] for r in res]), [np.asarray(r["segments"]).reshape(-1, 2) for r in res]))
    eng = seg._engine
    eng.profile(True)
    eng.profile_read()
    for _ in range(10):
        seg(wav=rows, in_second=False)
    torch.cuda.synchronize()
    prof = {k: v[0] / max(v[1], 1) for k, v in eng.profile_read().items() if v[1]}
    eng.profile(False)
    same = all(np.array_equal(runs[0][0], r[0]) for r in runs[1:])
    np.save(out + ".npy", runs[0][0])
    json.dump({"deterministic": bool(same), "segments": [s.tolist() for s in runs[0][1]], "stage_ms": prof},
              open(out + ".json", "w"))


def main():
    import numpy as np
    res = {}
    for flag in ("0", "1"):
        out = f"/tmp/streamk_{flag}"
        env = dict(os.environ, SYL_STREAMK=flag)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", out], env=env, timeout=300)
        if p.returncode != 0:
            print(f"SYL_STREAMK={flag}: child failed with exit code {p.returncode}")
            sys.exit(1)
        res[flag] = (np.load(out + ".npy"), json.load(open(out + ".json")))
    a, b = res["0"][0].astype(np.float64), res["1"][0].astype(np.float64)
    rel = float(np.linalg.norm(a - b) / np.linalg.norm(a))
    moved = sum(x != y for x, y in zip(res["0"][1]["segments"], res["1"][1]["segments"]))
    print("deterministic:", res["0"][1]["deterministic"], res["1"][1]["deterministic"])
    print("hidden_states rel diff stream-K vs data-parallel: %.3e" % rel)
    print("utterances with a moved boundary:", moved, "of", len(res["0"][1]["segments"]))
    for k in res["0"][1]["stage_ms"]:
        print("  %-14s %.4f -> %.4f ms" % (k, res["0"][1]["stage_ms"][k], res["1"][1]["stage_ms"].get(k, float("nan"))))
    ok = res["1"][1]["deterministic"] and rel < 1e-5
    print("OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "child":
        child(sys.argv[2])
    else:
        main()
