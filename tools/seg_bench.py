"""Segmentation stage alone on realistic states: one forward of the bench workload (speech-like synthetic weights),
then syl_segment on its hidden states, timed with CUDA events; also the ncu target for segment_kernel.
    python tools/seg_bench.py [B=32] [seconds=10] [reps=20]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
seg = Segmenter(model_ckpt=None, state_dict=syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM), device="cuda:0")
eng = seg._engine
g = torch.Generator().manual_seed(1)
wav = torch.randn(B, secs * 16000, generator=g).cuda()
n = torch.full((B,), secs * 16000, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
hid, sg, cnt, feat = eng.forward(wav, n, 2.6, 0.8)
hid = hid.clone(); torch.cuda.synchronize()
print("frames", hid.shape[1], "segments per clip", float(cnt.float().mean()))
for _ in range(3):
    eng.segment_states(hid, 2.6, 0.8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.segment_states(hid, 2.6, 0.8)
e1.record(); torch.cuda.synchronize()
print("segmentation stage (sqnorm + scan + pool): %.4f ms" % (e0.elapsed_time(e1) / reps))
