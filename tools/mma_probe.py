import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sylber_b200 import _lib
import gpu_util as G
lib = _lib.load_diag_library()      # diagnostic build (-DSYL_DIAG): the product library has no trace / probe entry points
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for ctas in (1, 148):
    for n in (16, 32, 48, 64, 96, 128, 256):
        iters = 2000
        lib.syl_mma_probe(n, iters, ctas, G.ptr(out), G.stream()); torch.cuda.synchronize()
        lib.syl_mma_probe(n, iters, ctas, G.ptr(out), G.stream()); torch.cuda.synchronize()
        cyc = int(out.item())
        print(f"ctas {ctas:3d}  N={n:3d}: {cyc/(iters*4):7.1f} cycles per MMA (ideal tensor time {128*n/256:.0f})")
