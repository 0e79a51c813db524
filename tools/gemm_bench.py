"""Stand-alone timing of the tensor-core GEMM through the C ABI (syl_gemm_f32) for the Segmenter's shapes."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sylber_b200 import _lib
import gpu_util as G

lib = _lib.load_library()
dev = torch.device("cuda", 0)
shapes = [(15968, 768, 3072, 1), (15968, 3072, 768, 1), (15968, 2304, 768, 1), (15968, 768, 768, 1), (63996, 512, 1536, 3)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (M, N, K, n_pass) in shapes:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.05
    need = lib.syl_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=dev); out = torch.empty(M, N, device=dev)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    def run():
        rc = lib.syl_gemm_f32(G.ptr(A), G.ptr(W), None, None, G.ptr(out), M, N, K, n_pass, 0, G.ptr(ws), need, G.stream())
        assert rc == 0, lib.syl_last_error(None)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n          # includes the two fp32->fp16 split kernels of the test entry point
    print(f"M{M} N{N} K{K} pass{n_pass}: {ms:.3f} ms/call (incl. operand split), {2*M*N*K/ms/1e9:.0f} TFLOP/s algorithmic")
