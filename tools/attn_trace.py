"""Timeline of attention7_kernel's CTA 0 (syl_attention_trace): per-unit latencies of the softmax chains and the MMA thread."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sylber_b200 import _lib
import gpu_util as G
lib = _lib.load_diag_library()      # diagnostic build (-DSYL_DIAG): the product library has no trace / probe entry points
dev = torch.device("cuda", 0)
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 499)
CAP = 4096
qkv = (torch.randn(B * T, 2304, device=dev) * 0.5).half()
out = torch.zeros(B * T, 768, dtype=torch.float16, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(3):
    lib.syl_attention(G.ptr(qkv), None, B, T, G.ptr(out), G.stream())
tr = torch.zeros(7, CAP, dtype=torch.int64, device=dev)
rc = lib.syl_attention_trace(G.ptr(qkv), None, B, T, G.ptr(out), G.ptr(tr), CAP, G.stream())
assert rc == 0
torch.cuda.synchronize()
tr = tr.cpu().numpy()
recs = []
for w in range(7):
    for v in tr[w]:
        if v == 0:
            break
        recs.append((int(v & 0xffffffffff), w, int(v >> 56) & 0xff, int(v >> 52) & 0xf, int(v >> 40) & 0xfff))
recs.sort()
t0 = recs[0][0]
names = {1: "wait_S", 2: "S_ready", 3: "pass_done", 6: "arrived", 7: "O_done", 8: "epi_done", 4: "mma_issue", 5: "mma_issued"}
if "-v" in sys.argv:
    for t, w, k, x, u in recs[:400]:
        print(f"{t - t0:8d}  {'MMA%d' % (w - 5) if w >= 5 else 'WG%d' % (w - 1)}  {names[k]:10s} x={x} u={u}")
# per-tile statistics
def stat(a):
    a = np.asarray(a, dtype=np.float64)
    return f"n={len(a):4d} mean={a.mean():7.0f} p50={np.median(a):7.0f} max={a.max():7.0f}" if len(a) else "n=0"
by = {}
for t, w, k, x, u in recs:
    by.setdefault((w, k), []).append((t, x, u))
for x in range(4):
    w = 1 + x
    waitS, ready, done, arr = (by.get((w, k), []) for k in (1, 2, 3, 6))
    n = min(len(waitS), len(ready), len(done), len(arr))
    print(f"tile {x}: wait for S   {stat([ready[i][0] - waitS[i][0] for i in range(n)])}")
    print(f"        softmax pass {stat([done[i][0] - ready[i][0] for i in range(n)])}")
    print(f"        fence+arrive {stat([arr[i][0] - done[i][0] for i in range(n)])}")
    print(f"        unit period  {stat([ready[i + 1][0] - ready[i][0] for i in range(n - 1)])}")
iss = sorted(sum((by.get((5 + x, 4), []) for x in range(2)), []))
isd = sorted(sum((by.get((5 + x, 5), []) for x in range(2)), []))
n = min(len(iss), len(isd))
print("MMA thread: issue step ", stat([isd[i][0] - iss[i][0] for i in range(n)]))
print("MMA thread: gap between steps", stat([iss[i + 1][0] - isd[i][0] for i in range(n - 1)]))
# latency from a warpgroup's arrive to the MMA thread starting that tile's next step, and from issue to S_ready
arr_map = {(x, u): t for (t, x, u) in sum((by.get((1 + x, 6), []) for x in range(4)), [])}
lat1, lat2 = [], []
items_seen = {}
# unit counters restart per item; match in time order instead
import bisect
for x in range(4):
    arr = [t for (t, _, _) in by.get((1 + x, 6), [])]
    steps = [(t, u) for (t, xx, u) in iss if xx == x and u > 0]
    ends = [(t, u) for (t, xx, u) in isd if xx == x]
    ready = [t for (t, _, _) in by.get((1 + x, 2), [])]
    for (t, u) in steps:
        i = bisect.bisect_right(arr, t) - 1
        if i >= 0:
            lat1.append(t - arr[i])
    for (t, u) in ends:
        i = bisect.bisect_left(ready, t)
        if i < len(ready):
            lat2.append(ready[i] - t)
print("arrive -> MMA thread starts the step ", stat(lat1))
print("step issued -> S ready in the warpgroup", stat(lat2))
print("total cycles", recs[-1][0] - t0)
