import sys, time, torch
sys.path.insert(0, ".")
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict
sd = syllabic_test_state_dict(9, 0)
g = torch.Generator().manual_seed(1)
wav = torch.randn(32, 160000, generator=g)
wl = [wav[i:i+1] for i in range(32)]
for ns in (1, 2, 3, 4):
    seg = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", streams=ns)
    for _ in range(4): seg(wav=wl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): r = seg(wav=wl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("streams", ns, "e2e ms", round(dt * 1e3, 3), "frames/s", round(32 * 499 / dt), flush=True)
    del seg
