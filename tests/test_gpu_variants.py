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
                                  ["SYL_STREAMK=1", "SYL_STREAMK_PCT=0"], ["SYL_STREAMK=1", "SYL_RESID_EPI=2"]])
def test_variant_matches_default_build(cuda, args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variant_check.py")] + args, capture_output=True, text=True,
                       timeout=800)
    print(p.stdout[-4000:])
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
