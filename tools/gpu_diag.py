"""First-contact diagnostics on the GPU box: exercise each kernel through the C ABI and print errors.
Not a test - prints numbers so that a failing kernel can be localised from one gpurun call."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sylber_b200 import _lib  # noqa: E402
import gpu_util as G  # noqa: E402


def section(name):
    print(f"\n=== {name} ===", flush=True)


def main():
    lib = _lib.load_library()
    dev = torch.device("cuda", 0)
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    torch.manual_seed(0)

    section("powf emulation")
    try:
        from oracle import segment_ref as R
        x = (torch.rand(1 << 20) * 2000 + 1e-3).float()
        y = torch.empty_like(x, device=dev)
        xd = x.to(dev)
        rc = lib.syl_powf_half(G.ptr(xd), G.ptr(y), x.numel(), G.stream())
        torch.cuda.synchronize()
        ref = np.array([R._lib().syl_oracle_powf_half(float(v)) for v in x[:200000].numpy()], dtype=np.float32)
        got = y.cpu().numpy()[:200000]
        print("rc", rc, "mismatches vs libm powf:", int((ref != got).sum()), "of", len(ref),
              "| vs sqrt:", int((np.sqrt(x[:200000].numpy()) != got).sum()))
    except Exception:
        traceback.print_exc()

    section("gemm")
    for (M, N, K) in [(128, 256, 64), (300, 256, 128), (1000, 768, 768), (4096, 2304, 768), (15968, 768, 3072)]:
        try:
            A = torch.randn(M, K, device=dev)
            W = torch.randn(N, K, device=dev) * 0.05
            bias = torch.randn(N, device=dev)
            ref = A.double() @ W.double().t() + bias.double()
            for n_pass in (1, 3):
                out = G.gemm_f32(lib, A, W, bias=bias, n_pass=n_pass)
                print(f"M{M} N{N} K{K} pass{n_pass}: rel {G.rel_err(out, ref):.3e} max {G.max_rel(out, ref):.3e}", flush=True)
            res = torch.randn(M, N, device=dev)
            out = G.gemm_f32(lib, A, W, bias=bias, residual=res, n_pass=3, act=1)
            ref2 = torch.nn.functional.gelu((A.double() @ W.double().t() + bias.double())) + res.double()
            print(f"   gelu+residual: rel {G.rel_err(out, ref2):.3e}")
        except Exception:
            traceback.print_exc()

    section("attention")
    try:
        import ctypes
        for (B, T, lens) in [(2, 499, None), (3, 143, [143, 100, 17]), (1, 1000, [777])]:
            qkv = torch.randn(B * T, 2304, device=dev)
            qkv[:, :768] *= 0.125 * 1.5
            q16 = qkv.half()
            out = torch.zeros(B * T, 768, dtype=torch.float16, device=dev)
            kv = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=dev)
            rc = lib.syl_attention(G.ptr(q16), G.ptr(kv), B, T, G.ptr(out), G.stream())
            torch.cuda.synchronize()
            x = q16.float().view(B, T, 3, 12, 64)
            q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
            s = q @ k.transpose(2, 3)
            if lens is not None:
                m = torch.arange(T, device=dev)[None, :] >= kv[:, None]
                s = s.masked_fill(m[:, None, None, :], float("-inf"))
            ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, 768)
            print(f"B{B} T{T} lens{lens}: rc {rc} rel {G.rel_err(out.float(), ref):.3e} max {G.max_rel(out.float(), ref):.3e}", flush=True)
    except Exception:
        traceback.print_exc()

    section("segmentation")
    try:
        from oracle import segment_ref as R
        rng = np.random.default_rng(0)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from seg_cases import plateau_states
        bad = 0
        for trial in range(6):
            B, T = 8, int(rng.integers(20, 500))
            st = np.stack([plateau_states(rng, T) for _ in range(B)])
            seg, cnt, feat = G.segment(lib, torch.from_numpy(st).to(dev))
            for b in range(B):
                ref = R.c_get_segment(st[b], 2.6, 0.8)
                got = seg[b, :cnt[b]].astype(np.int64)
                ok = len(ref) == len(got) and np.array_equal(ref, got)
                if ok and len(ref):
                    ok = np.array_equal(R.c_segment_mean(st[b], ref), feat[b, :cnt[b]], equal_nan=True)
                bad += (not ok)
        print("segmentation mismatching utterances:", bad, "of 48")
    except Exception:
        traceback.print_exc()

    section("model stages vs oracle")
    try:
        from oracle.hubert_ref import hubert_forward, num_frames
        from sylber_b200.weights import syllabic_test_state_dict
        from sylber_b200.segmenter import _Engine
        sd = syllabic_test_state_dict(9, 0)
        g = torch.Generator().manual_seed(1)
        B, n = 3, 48000
        wav = torch.randn(B, n, generator=g)
        lens = [48000, 30000, 16123]
        for b, l in enumerate(lens):
            wav[b, l:] = 0
        stages = {}
        t0 = time.time()
        ref = hubert_forward(sd, wav, lens, 9, stages=stages)
        print("oracle forward s:", round(time.time() - t0, 2))
        for mode in ("parity", "strict", "fast", "exact"):
            eng = _Engine(sd, 9, "cuda:0", mode)
            hidden, seg, cnt, feat = eng.forward(wav.to(dev), torch.tensor(lens, dtype=torch.int32, device=dev), 2.6, 0.8)
            torch.cuda.synchronize()
            T = hidden.shape[1]
            print(f"[{mode}] hidden rel {G.rel_err(hidden.cpu(), ref):.3e} max {G.max_rel(hidden.cpu(), ref):.3e}  segs {cnt.tolist()}")
            if mode == "parity":
                for i in range(7):
                    L = stages[f"conv{i}"].shape[2]
                    got = eng.read_stage(f"conv{i}", (B, L, 512)).cpu()
                    print(f"   conv{i}: rel {G.rel_err(got, stages[f'conv{i}'].transpose(1, 2)):.3e}")
                got = eng.read_stage("pos", (B, T, 768)).cpu()
                print(f"   pos: rel {G.rel_err(got, stages['pos']):.3e}")
                for nl in (0, 1, 2, 5):
                    eng.set_active_layers(nl)
                    hid, _, _, _ = eng.forward(wav.to(dev), torch.tensor(lens, dtype=torch.int32, device=dev), 2.6, 0.8, segment=False)
                    key = "enc_in" if nl == 0 else f"layer{nl - 1}"
                    print(f"   after {nl} layers: rel {G.rel_err(hid.cpu(), stages[key]):.3e}")
                eng.set_active_layers(-1)
            del eng
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
