"""Error budget of the split-precision sites (run on the GPU box: python tests/mode_sweep.py).

For a set of precision masks (include/sylber_b200.h SYL_SPLIT_*) prints the relative Frobenius error of the final
hidden states against the fp32 CPU oracle on a padded 2-clip batch, and the device time of the bench workload
(batch 32 x 10 s, CUDA-graph replay).  Test infrastructure: this is the one place outside the tests that calls oracle/."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict
from oracle.hubert_ref import hubert_forward

C1, ENC, C2, C3, C4, C5, C6, FPROJ, POS = 8, 4, 16, 32, 64, 128, 256, 512, 1024
MASKS = {
    "fast (none)": 0,
    "c6 fproj": C6 | FPROJ,
    "c5-6 fproj": C5 | C6 | FPROJ,
    "c4-6 fproj": C4 | C5 | C6 | FPROJ,
    "c4-6 fproj pos": C4 | C5 | C6 | FPROJ | POS,
    "c3-6 fproj": C3 | C4 | C5 | C6 | FPROJ,
    "c3-6 fproj pos": C3 | C4 | C5 | C6 | FPROJ | POS,
    "c2-6 fproj": C2 | C3 | C4 | C5 | C6 | FPROJ,
    "parity (c2-6 fproj pos)": C2 | C3 | C4 | C5 | C6 | FPROJ | POS,
    "strict (+c1)": C1 | C2 | C3 | C4 | C5 | C6 | FPROJ | POS,
    "c1 only": C1,
    "pos only": POS,
}
sd = syllabic_test_state_dict(9, 0)
gen = torch.Generator().manual_seed(1)
lens = [48000, 30000]
batch = torch.zeros(2, 48000)
for i, n in enumerate(lens):
    batch[i, :n] = torch.randn(n, generator=gen)
ref = hubert_forward(sd, batch, lens, 9).numpy()
g2 = torch.Generator().manual_seed(1)
big = torch.randn(32, 160000, generator=g2)
for name, mask in MASKS.items():
    s = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", mode=mask)
    eng = s._engine
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    hid, _, _, _ = eng.forward(batch.to(eng.device), torch.tensor(lens, dtype=torch.int32, device=eng.device), 2.6, 0.8, segment=False)
    got = hid.cpu().numpy()
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    wav = big.to(eng.device); n = torch.full((32,), 160000, dtype=torch.int32, device=eng.device)
    for _ in range(3):
        eng.forward(wav, n, 2.6, 0.8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward(wav, n, 2.6, 0.8)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} mask {mask:5d}  rel err {err:.3e}  {e0.elapsed_time(e1) / 10:.3f} ms/step", flush=True)
    del s, eng
    torch.cuda.empty_cache()
