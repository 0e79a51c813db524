"""End-to-end time of Segmenter.__call__ (32 x 10 s, pinned inputs) for different sub-batch splits."""
import sys, time, torch
sys.path.insert(0, ".")
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
g = torch.Generator().manual_seed(1)
wav = torch.randn(32, 160000, generator=g).pin_memory()
wl = [wav[i:i+1] for i in range(32)]
for split in ([8, 12, 12], [12, 12, 8], [10, 12, 10], [12, 20], [16, 16], [8, 24], [8, 12, 12]):
    seg = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", streams=len(split), sub_batch_sizes=split, max_batch=32)
    for _ in range(4): seg(wav=wl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): r = seg(wav=wl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("split", split, "e2e ms", round(dt * 1e3, 3), "frames/s", round(32 * 499 / dt), flush=True)
    del seg
