"""Two eager (graph-off) forwards of the bench workload, for ncu: the second forward's launches are the warm ones.
    ncu --set full --import-source on --clock-control none -s 83 -c 83 -o out python tools/ncu_forward.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
seg = Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM), device="cuda:0", mode=mode)
eng = seg._engine
eng.lib.syl_set_graph_mode(eng.handle, 0)
g = torch.Generator().manual_seed(1)
wav = torch.randn(32, 160000, generator=g).cuda()
n = torch.full((32,), 160000, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(2):
    eng.forward(wav, n, 2.6, 0.8)
torch.cuda.synchronize()
print("launches per forward:", eng.launch_count(True))
