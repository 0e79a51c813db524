"""Per-stage relative error against the fp32 CPU oracle for one precision preset (default exact), on a padded 2-clip
batch: conv0..conv6, positional conv, and the hidden states after 0..9 encoder layers.  Test infrastructure (uses oracle/).
    python tools/stage_drift.py [mode]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
from oracle.hubert_ref import hubert_forward
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
gen = torch.Generator().manual_seed(1)
lens = [48000, 30000]
batch = torch.zeros(2, 48000)
for i, n in enumerate(lens):
    batch[i, :n] = torch.randn(n, generator=gen)
stages = {}
ref = hubert_forward(sd, batch, lens, 9, stages=stages).numpy()
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
s = Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", mode=mode)
eng = s._engine
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
n_dev = torch.tensor(lens, dtype=torch.int32, device=eng.device)
hid, _, _, _ = eng.forward(batch.to(eng.device), n_dev, 2.6, 0.8, segment=False)
print(mode, "final", "%.3e" % rel(hid.cpu().numpy(), ref))
for i in range(7):
    L = stages[f"conv{i}"].shape[2]
    got = eng.read_stage(f"conv{i}", (2, L, 512)).cpu().numpy()
    print(f"conv{i}", "%.3e" % rel(got, stages[f"conv{i}"].transpose(1, 2).numpy()))
T = hid.shape[1]
print("pos", "%.3e" % rel(eng.read_stage("pos", (2, T, 768)).cpu().numpy(), stages["pos"].numpy()))
for n in range(0, 10):
    eng.set_active_layers(n)
    h, _, _, _ = eng.forward(batch.to(eng.device), n_dev, 2.6, 0.8, segment=False)
    want = stages["enc_in"] if n == 0 else stages[f"layer{n - 1}"]
    print(f"after {n} layers", "%.3e" % rel(h.cpu().numpy(), want.numpy()))
eng.set_active_layers(-1)
