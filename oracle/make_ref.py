"""Installs the UNMODIFIED reference package into the git-ignored oracle/_ref/ (TEST INFRASTRUCTURE recipe).

    python oracle/make_ref.py            # needs /root/reference; a no-op message when it is absent

What it does: `pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>`
(the offline install the bench contract names; the copy under /tmp is needed because the checkout is read-only and
setuptools writes build/ and *.egg-info next to setup.py).  Nothing is copied into tracked paths: oracle/_ref/ is
listed in .gitignore but not in .gpurunignore, so the installed package travels to the GPU box where
`bench.py --impl reference` imports it (`load_reference()` below) and times `sylber.Segmenter(device='cpu')` -
the reference's own code (sylber/model/sylber.py:28-138, sylber/utils/segment_utils.py:68-131) on the box's cores.

The reference's optional dependencies torchode / torchdiffeq / gateloop_transformer / vector_quantize_pytorch /
lightning are not in this image and are never touched by the Segmenter path; `load_reference()` puts empty stub
modules in sys.modules for them before the import, exactly as tests/golden/make_golden.py does.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")


def install(force=False):
    if os.path.isdir(os.path.join(REF_DST, "sylber")) and not force:
        return REF_DST
    if not os.path.isdir(REF_SRC):
        print("oracle/make_ref.py: /root/reference is absent; nothing installed")
        return None
    shutil.rmtree(REF_DST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "docs", "*.ipynb", "samples"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", REF_DST, src]
        subprocess.check_call(cmd)
    assert os.path.isfile(os.path.join(REF_DST, "sylber", "model", "sylber.py"))
    return REF_DST


def available():
    return os.path.isfile(os.path.join(REF_DST, "sylber", "model", "sylber.py"))


def load_reference():
    """Import the installed reference (oracle/_ref/sylber) with the uninstalled optional modules stubbed.
    Returns (sylber module, get_segment)."""
    import torch
    if not available():
        raise ImportError("oracle/_ref is empty: run python oracle/make_ref.py where /root/reference exists")
    stubs = {"torchode": ["Tsit5"], "torchdiffeq": ["odeint"], "gateloop_transformer": ["SimpleGateLoopLayer"],
             "vector_quantize_pytorch": ["GroupedResidualVQ"], "lightning": []}
    for name, attrs in stubs.items():
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    if not hasattr(sys.modules["lightning"], "LightningModule"):
        sys.modules["lightning"].LightningModule = torch.nn.Module
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import sylber
    from sylber.utils.segment_utils import get_segment
    assert os.path.realpath(sylber.__file__).startswith(os.path.realpath(REF_DST)), sylber.__file__
    return sylber, get_segment


def reference_segmenter(state_dict, n_layers=9):
    """`sylber.Segmenter(model_ckpt=None, device='cpu')` with `state_dict` loaded into its HubertModel.  The config
    directory stands in for the hub id `facebook/hubert-base-ls960` (HubertConfig() defaults are that config)."""
    from transformers import HubertConfig
    sylber, _ = load_reference()
    with tempfile.TemporaryDirectory() as d:
        HubertConfig().save_pretrained(d)
        seg = sylber.Segmenter(model_ckpt=None, speech_upstream=d, encoding_layer=n_layers, device="cpu")
    missing, unexpected = seg.speech_model.load_state_dict(state_dict, strict=False)
    assert set(missing) <= {"masked_spec_embed"} and not unexpected, (missing, unexpected)
    return seg


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
