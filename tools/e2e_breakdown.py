"""Where does the end-to-end Segmenter.__call__ time go (host staging, H2D, device, D2H)?  Diagnostic only."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict

seg = Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0), device="cuda:0")
eng = seg._engine
g = torch.Generator().manual_seed(1)
wav = torch.randn(32, 160000, generator=g)
wl = [wav[i:i + 1] for i in range(32)]
for _ in range(3):
    seg(wav=wl)
def t(f, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("full __call__ ms", round(t(lambda: seg(wav=wl)), 3))
host = torch.empty((32, 160000), dtype=torch.float32, pin_memory=True)
def stage():
    for i, r in enumerate(wl): host[i].copy_(r[0])
print("host staging (32 row copies into pinned) ms", round(t(stage), 3))
print("pinned alloc 20MB ms", round(t(lambda: torch.empty((32, 160000), dtype=torch.float32, pin_memory=True)), 3))
print("H2D 20MB ms", round(t(lambda: host.to("cuda", non_blocking=True)), 3))
wd = host.to("cuda"); nd = torch.full((32,), 160000, dtype=torch.int32, device="cuda")
print("device forward ms", round(t(lambda: eng.forward(wd, nd, 2.6, 0.8)), 3))
hid, sg, cnt, feat = eng.forward(wd, nd, 2.6, 0.8)
hp = torch.empty(hid.shape, dtype=torch.float32, pin_memory=True)
print("D2H hidden 49MB (pinned, reused) ms", round(t(lambda: hp.copy_(hid, non_blocking=True)), 3))
print("pinned alloc 49MB ms", round(t(lambda: torch.empty(hid.shape, dtype=torch.float32, pin_memory=True)), 3))
print("cnt.cpu ms", round(t(lambda: cnt.cpu()), 3))
nm = int(cnt.max())
print("feat slice D2H ms", round(t(lambda: feat[:, :nm].cpu()), 3), "n_max", nm)
