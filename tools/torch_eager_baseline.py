"""The GPU-library bar: transformers' HubertModel (cuDNN convs, cuBLAS linears, SDPA attention) in
PyTorch eager on one B200, same workload as bench.py (batch 32 x 10 s, 9 layers, random weights), model forward only
(no segmentation).  Prints one JSON line per precision setting.  Diagnostic: nothing here is on the product path."""
import json, sys, time
import torch
from transformers import HubertConfig, HubertModel

B, N, LAYERS = 32, 160000, 9
torch.manual_seed(0)
model = HubertModel(HubertConfig(num_hidden_layers=LAYERS)).eval().cuda()
wav = torch.randn(B, N, generator=torch.Generator().manual_seed(1)).cuda()
T = model._get_feat_extract_output_lengths(N)
T = int(T)

def run(name, ctx, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad(), ctx():
        for _ in range(3):
            model(wav)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = model(wav).last_hidden_state
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"impl": "torch eager HubertModel", "precision": name, "ms_per_step": round(ms, 3),
                      "frames_per_s": round(B * T / ms * 1e3), "torch": torch.__version__}), flush=True)

import contextlib
run("fp32 (tf32 off)", contextlib.nullcontext, False)
run("tf32", contextlib.nullcontext, True)
run("fp16 autocast", lambda: torch.autocast("cuda", dtype=torch.float16), True)
run("bf16 autocast", lambda: torch.autocast("cuda", dtype=torch.bfloat16), True)
