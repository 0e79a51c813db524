"""A/B check of a kernel variant that is selected by environment variables, on one GPU.

    python tools/variant_check.py SYL_STREAMK=1                  # stream-K GEMM schedule (csrc/gemm3_tc.cuh)
    python tools/variant_check.py SYL_RESID_EPI=2 --exact        # residual add in the GEMM epilogue: must be bit-identical
    python tools/variant_check.py SYL_STREAMK=1 SYL_RESID_EPI=1
    python tools/variant_check.py child OUT                      # (internal) one process, variant taken from its environment

The parent runs two children under a timeout - the baseline (variables unset) and the variant.  Each runs batch
32 x 10 s (189 tiles for the N = 768 GEMMs, the case stream-K is for) three times - eager launch, then CUDA-graph
replays - and saves hidden states and segments.  Required: the variant's runs are bit-identical to each other (stream-K
adds its partial sums in a fixed order), hidden states within 1e-5 relative of the baseline (or identical with
--exact); reported: how many utterances changed a segment boundary (a different summation order can move a threshold
decision; the count should be zero or close to it) and the per-stage device times of both."""
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
    prof = {k: v[0] / 10 for k, v in eng.profile_read().items() if v[1]}       # ms per forward
    eng.profile(False)
    same = all(np.array_equal(runs[0][0], r[0]) for r in runs[1:])
    np.save(out + ".npy", runs[0][0])
    json.dump({"deterministic": bool(same), "segments": [s.tolist() for s in runs[0][1]], "stage_ms": prof},
              open(out + ".json", "w"))


def main():
    import numpy as np
    exact = "--exact" in sys.argv
    variant = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    if not variant:
        print(__doc__)
        sys.exit(2)
    res = {}
    for name, extra in (("base", {}), ("variant", variant)):
        out = f"/tmp/variant_check_{name}"
        env = {k: v for k, v in os.environ.items() if k not in variant}
        env.update(extra)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child", out], env=env, timeout=300)
        if p.returncode != 0:
            print(f"{name} {extra}: child failed with exit code {p.returncode}")
            sys.exit(1)
        res[name] = (np.load(out + ".npy"), json.load(open(out + ".json")))
    a, b = res["base"][0].astype(np.float64), res["variant"][0].astype(np.float64)
    rel = float(np.linalg.norm(a - b) / np.linalg.norm(a))
    moved = sum(x != y for x, y in zip(res["base"][1]["segments"], res["variant"][1]["segments"]))
    print("variant:", variant)
    print("deterministic (base, variant):", res["base"][1]["deterministic"], res["variant"][1]["deterministic"])
    print("hidden_states rel diff variant vs base: %.3e" % rel, "(identical)" if np.array_equal(res["base"][0], res["variant"][0]) else "")
    print("utterances with a moved boundary:", moved, "of", len(res["base"][1]["segments"]))
    total = [0.0, 0.0]
    for k in res["base"][1]["stage_ms"]:
        x, y = res["base"][1]["stage_ms"][k], res["variant"][1]["stage_ms"].get(k, float("nan"))
        total[0] += x
        total[1] += y
        print("  %-14s %.4f -> %.4f ms" % (k, x, y))
    print("  %-14s %.4f -> %.4f ms" % ("sum", total[0], total[1]))
    ok = res["variant"][1]["deterministic"] and (np.array_equal(res["base"][0], res["variant"][0]) if exact else rel < 1e-5)
    print("OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "child":
        child(sys.argv[2])
    else:
        main()
