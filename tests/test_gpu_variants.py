"""Kernel variants that are selected by environment variables and are NOT the default path.
They were written after the GPU budget of the round that produced them was spent, so these checks are opt-in
(SYL_TEST_VARIANTS=1) until a variant has been run and timed on a B200; each one runs tools/variant_check.py, which
compares a child process running the variant with one running the default build on batch 32 x 10 s."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SYL_TEST_VARIANTS") != "1", reason="opt-in: set SYL_TEST_VARIANTS=1")]


@pytest.mark.parametrize("args", [["SYL_RESID_EPI=1", "--exact"], ["SYL_RESID_EPI=2", "--exact"], ["SYL_STREAMK=1"],
                                  ["SYL_STREAMK=1", "SYL_STREAMK_PCT=0"], ["SYL_STREAMK=1", "SYL_RESID_EPI=2"], ["SYL_CONV0_MB=5", "--exact"], ["SYL_LN_WARPS=4", "--exact"]])
def test_variant_matches_default_build(cuda, args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variant_check.py")] + args, capture_output=True, text=True,
                       timeout=800)
    print(p.stdout[-4000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


_GEMM_SK = r"""
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import gpu_util as G
from sylber_b200 import _lib
lib = _lib.load_library()
cuda = torch.device("cuda", 0)
for M, N, K, n_pass in [(15968, 768, 768, 1), (15968, 768, 3072, 1), (15968, 768, 512, 3), (4096, 2304, 768, 1), (20000, 256, 64, 1),
                        (19999, 512, 1536, 3)]:
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=cuda)
    W = torch.randn(N, K, device=cuda) * 0.05
    bias = torch.randn(N, device=cuda)
    ref = A.double() @ W.double().t() + bias.double()
    outs = [G.gemm_f32(lib, A, W, bias=bias, n_pass=n_pass) for _ in range(3)]
    torch.cuda.synchronize()
    err = G.rel_err(outs[0], ref)
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    worst = float((outs[0].double() - ref).abs().max())
    print(M, N, K, n_pass, "rel %%.3e worst %%.3e deterministic %%s" %% (err, worst, same), flush=True)
    assert same and err < (6e-4 if n_pass == 1 else 4e-5) and worst < (0.05 if n_pass == 1 else 0.005)
print("OK")
"""


def test_streamk_gemm_vs_fp64(cuda):
    """The stream-K schedule on single GEMMs against fp64 (every launch here has more tiles than clusters; PCT=0 forces
    the schedule even when the last round is well filled), three runs each: a wrong fix-up shows as a large worst-element
    error in the tiles that were cut, a race as run-to-run differences."""
    env = dict(os.environ, SYL_STREAMK="1", SYL_STREAMK_PCT="0")
    code = _GEMM_SK % (ROOT, os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    print(p.stdout[-3000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
