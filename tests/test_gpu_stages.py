"""GPU parity of the per-stage C-ABI entry points (include/sylber_b200.h: syl_conv_frontend, syl_encoder_layer,
syl_read_stage) against the fp32 CPU oracle - the same entry points ncu is pointed at."""
import ctypes

import numpy as np
import pytest
import torch

import gpu_util as G
from oracle.hubert_ref import feature_encoder, encoder_layer, num_frames
from sylber_b200 import Segmenter
from sylber_b200.weights import syllabic_test_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    sd = syllabic_test_state_dict(9, 0)
    return sd, Segmenter(model_ckpt=None, state_dict=sd, device="cuda:0")


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_conv_frontend_vs_oracle(setup):
    sd, seg = setup
    eng = seg._engine
    g = torch.Generator().manual_seed(21)
    wav = torch.randn(3, 24000, generator=g)
    T = num_frames(24000)
    stages = {}
    want = feature_encoder(sd, wav, stages=stages)              # (B, 512, T)
    need = int(eng.lib.syl_workspace_bytes(eng.handle, 3, 24000))
    ws = torch.empty(need, dtype=torch.uint8, device=eng.device)
    feats = torch.empty((3, T, 512), dtype=torch.float32, device=eng.device)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    rc = eng.lib.syl_conv_frontend(eng.handle, G.ptr(wav.to(eng.device)), 3, 24000, G.ptr(feats), G.ptr(ws), need, G.stream())
    assert rc == 0, eng.lib.syl_last_error(eng.handle)
    torch.cuda.synchronize()
    assert _rel(feats.cpu().numpy(), want.transpose(1, 2).numpy()) < 1e-3
    # intermediate activations through syl_read_stage: conv0 is fp32 math rounded to fp16 (hi only in the default mode)
    L0 = stages["conv0"].shape[2]
    out = torch.empty((3, L0, 512), dtype=torch.float32, device=eng.device)
    rc = eng.lib.syl_read_stage(eng.handle, b"conv0", G.ptr(out), out.numel(), G.stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert _rel(out.cpu().numpy(), stages["conv0"].transpose(1, 2).numpy()) < 6e-4
    torch.cuda.set_stream(torch.cuda.default_stream())


@pytest.mark.parametrize("layer,T,valid", [(0, 143, None), (4, 300, [300, 211, 17]), (8, 499, [499, 1, 250])])
def test_encoder_layer_vs_oracle(setup, layer, T, valid):
    sd, seg = setup
    eng = seg._engine
    B = 3
    g = torch.Generator().manual_seed(layer)
    h_in = torch.randn(B, T, 768, generator=g)
    key_bias = None
    vf = None
    if valid is not None:
        vt = torch.tensor(valid)
        mask = torch.arange(T)[None, :] < vt[:, None]
        key_bias = torch.zeros(B, 1, 1, T).masked_fill(~mask[:, None, None, :], float("-inf"))
        vf = vt.to(torch.int32).to(eng.device)
    want = encoder_layer(sd, layer, h_in, key_bias)
    n = T
    for k, s in zip((2, 2, 3, 3, 3, 3, 10), (2, 2, 2, 2, 2, 2, 5)):      # smallest sample count with exactly T frames
        n = (n - 1) * s + k
    need = int(eng.lib.syl_workspace_bytes(eng.handle, B, n))
    ws = torch.empty(need, dtype=torch.uint8, device=eng.device)
    h_out = torch.empty((B, T, 768), dtype=torch.float32, device=eng.device)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    rc = eng.lib.syl_encoder_layer(eng.handle, layer, G.ptr(h_in.to(eng.device)), G.ptr(vf), B, T, G.ptr(h_out), G.ptr(ws), need,
                                   G.stream())
    assert rc == 0, eng.lib.syl_last_error(eng.handle)
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    got = h_out.cpu().numpy()
    assert np.isfinite(got).all()
    assert _rel(got, want.numpy()) < 5e-4          # one single-pass fp16 layer: ~1e-4


def test_stage_entry_points_reject_bad_arguments(setup):
    _, seg = setup
    eng = seg._engine
    x = torch.zeros(8, device=eng.device)
    assert eng.lib.syl_encoder_layer(eng.handle, 99, G.ptr(x), None, 1, 1, G.ptr(x), G.ptr(x), 8, G.stream()) != 0
    assert b"bad arguments" in eng.lib.syl_last_error(eng.handle)
    assert eng.lib.syl_conv_frontend(eng.handle, G.ptr(x), 1, 100, G.ptr(x), G.ptr(x), 8, G.stream()) != 0
    need = int(eng.lib.syl_workspace_bytes(eng.handle, 2, 16000))
    assert eng.lib.syl_conv_frontend(eng.handle, G.ptr(x), 2, 16000, G.ptr(x), G.ptr(x), need - 1, G.stream()) == -4   # SYL_E_WORKSPACE


@pytest.mark.parametrize("n", [48000, 160000, 82000])
def test_positional_conv_stage_vs_oracle(setup, n):
    """The single-pass positional conv pairs taps into N = 96 MMAs and exchanges half results between neighbouring
    rows (posconv.cuh); checked here at tile boundaries (T = 149, 499, 255 + 1) against the oracle's pos stage."""
    from oracle.hubert_ref import hubert_forward
    sd, seg = setup
    eng = seg._engine
    g = torch.Generator().manual_seed(n)
    wav = torch.randn(2, n, generator=g)
    lens = [n, n - 5000]
    wav[1, lens[1]:] = 0
    stages = {}
    hubert_forward(sd, wav, lens, 9, stages=stages)
    T = num_frames(n)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    eng.forward(wav.to(eng.device), torch.tensor(lens, dtype=torch.int32, device=eng.device), 2.6, 0.8, segment=False, slot="stage_test")
    got = eng.read_stage("pos", (2, T, 768)).cpu().numpy()
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    want = stages["pos"].numpy()
    assert _rel(got, want) < 1e-3
    # row-wise: no frame may be off (a wrong neighbour exchange would corrupt single rows, not the norm)
    row_err = np.linalg.norm(got - want, axis=-1) / (np.linalg.norm(want, axis=-1) + 1e-6)
    assert row_err.max() < 5e-3, (row_err.argmax(), row_err.max())
