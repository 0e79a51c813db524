"""One attention launch shape for ncu: python tools/attn_one.py B T"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from sylber_b200 import _lib
import gpu_util as G
lib = _lib.load_library()
dev = torch.device("cuda", 0)
B, T = int(sys.argv[1]), int(sys.argv[2])
qkv = (torch.randn(B * T, 2304, device=dev) * 0.5).half()
out = torch.zeros(B * T, 768, dtype=torch.float16, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for _ in range(4):
    lib.syl_attention(G.ptr(qkv), None, B, T, G.ptr(out), G.stream())
torch.cuda.synchronize()
