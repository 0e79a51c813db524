"""Segment agreement with the fp32 CPU oracle over several batches of config 2 (32 x 10 s each, different audio seeds) and
presets - a larger sample than tests/test_gpu_agreement.py.  Test infrastructure (uses oracle/).
    python tools/agreement_sweep.py [n_batches=4]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM
from oracle.hubert_ref import hubert_forward
from oracle import agreement as A

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.set_num_threads(os.cpu_count())
out = {}
for bias in (SPEECH_LIKE_BIAS_NORM, 2.1):
    sd = syllabic_test_state_dict(9, 0, bias)
    segs = {m: Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0", mode=m) for m in ("fast", "parity", "exact")}
    tot = {m: [0, 0, 0] for m in segs}          # agree, clips, unexplained
    for seed in range(10, 10 + n_batches):
        wav = torch.randn(32, 160000, generator=torch.Generator().manual_seed(seed))
        ref = np.concatenate([hubert_forward(sd, wav[k:k + 8], [160000] * 8, 9).numpy() for k in range(0, 32, 8)])
        clips = [wav[i:i + 1] for i in range(32)]
        for m, s in segs.items():
            res = s(wav=clips, in_second=False)
            recs = [A.compare_utterance(ref[i], res[i]["hidden_states"], res[i]["segments"]) for i in range(32)]
            summ = A.summarize(recs)
            tot[m][0] += summ["agree"]; tot[m][1] += 32; tot[m][2] += sum(not f["explained"] for f in summ["flips"])
    out[f"bias_norm_{bias}"] = {m: {"identical": v[0], "clips": v[1], "unexplained_flips": v[2]} for m, v in tot.items()}
    print(bias, out[f"bias_norm_{bias}"], flush=True)
print(json.dumps(out))
