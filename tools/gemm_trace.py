"""In-kernel timeline of gemm3_tc_kernel (diagnostic build): where the time of ONE launch goes besides the MMAs.
CTA 0 stamps clock64 at its role boundaries (slots: csrc/gemm_tc.cuh); two back-to-back launches give the gap between
one kernel's exit and the next one's entry (globaltimer).    python tools/gemm_trace.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sylber_b200 import _lib
import gpu_util as G

lib = _lib.load_diag_library()
dev = torch.device("cuda", 0)
shapes = [("QKV", 15968, 2304, 768), ("out-proj", 15968, 768, 768), ("FFN1", 15968, 3072, 768), ("FFN2", 15968, 768, 3072)]
for name, M, N, K in shapes:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.05
    need = lib.syl_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=dev); out = torch.empty(M, N, device=dev)
    tr = [torch.zeros(128, dtype=torch.int64, device=dev) for _ in range(2)]
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    def run(t=None):
        lib.syl_gemm_set_trace(G.ptr(t))
        rc = lib.syl_gemm_f32(G.ptr(A), G.ptr(W), None, None, G.ptr(out), M, N, K, 1, 0, G.ptr(ws), need, G.stream())
        assert rc == 0, lib.syl_last_error(None)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(tr[0]); e1.record(); torch.cuda.synchronize()
    t = tr[0].cpu().tolist()
    c0 = t[1]
    rel = lambda k: t[k] - c0
    tiles = [i for i in range(16) if t[10 + 3 * i]]
    print(f"== {name}: M{M} N{N} K{K}, {len(tiles)} tiles on CTA 0; launch (events, incl. 2 operand-split kernels) {e0.elapsed_time(e1)*1e3:.1f} us; "
          f"kernel entry->exit {t[102]-t[0]} ns = {t[101]-c0} cycles ({(t[101]-c0)/max(t[102]-t[0],1):.2f} GHz)")
    print(f"   prologue done {rel(2)}   first TMA issued {rel(3)}   first operands landed {rel(9)}")
    for i in tiles:
        print(f"   tile {i}: acc free {rel(8+3*i)}  operands {rel(9+3*i)}  committed {rel(10+3*i)} (mainloop {t[10+3*i]-t[9+3*i]})"
              f" | epilogue: acc complete {rel(64+2*i)}  stores committed {rel(65+2*i)} ({t[65+2*i]-t[64+2*i]})")
    print(f"   stores complete {rel(100)}   exit {rel(101)}   tail after last commit {t[101]-t[10+3*tiles[-1]]} cycles")
lib.syl_gemm_set_trace(None)
