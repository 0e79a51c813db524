"""Stand-alone timing of attention_kernel through the C ABI (syl_attention)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sylber_b200 import _lib
import gpu_util as G
lib = _lib.load_diag_library() if os.environ.get("SYL_DIAG_LIB") == "1" else _lib.load_library()
dev = torch.device("cuda", 0)
for (B, T) in [(32, 499), (8, 2999)]:
    qkv = (torch.randn(B * T, 2304, device=dev) * 0.5).half()
    out = torch.zeros(B * T, 768, dtype=torch.float16, device=dev)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    run = lambda: lib.syl_attention(G.ptr(qkv), None, B, T, G.ptr(out), G.stream())
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4 * B * T * T * 768
    by = 4 * B * T * 768 * 2
    print(f"B{B} T{T}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s  {by/ms/1e6:.0f} GB/s algorithmic")
