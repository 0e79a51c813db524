"""oracle/_ref: the unmodified reference package installed by oracle/make_ref.py (what `bench.py --impl reference` and the
cpu_baseline leg time).  Checks that the installed copy is the reference's own code and that the oracle port used by
the parity tests agrees with it on a full Segmenter call."""
import filecmp
import os

import numpy as np
import pytest
import torch

from oracle import make_ref
from oracle import segment_ref
from oracle.hubert_ref import hubert_forward
from sylber_b200.weights import syllabic_test_state_dict, SPEECH_LIKE_BIAS_NORM

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not (os.path.isdir(REF) or make_ref.available()), reason="needs /root/reference or a built oracle/_ref")


def test_installed_reference_is_unmodified_and_matches_the_port():
    if os.path.isdir(REF):
        make_ref.install()
        for rel in ("sylber/model/sylber.py", "sylber/utils/segment_utils.py"):
            assert filecmp.cmp(os.path.join(REF, rel), os.path.join(make_ref.REF_DST, rel), shallow=False), rel
    sd = syllabic_test_state_dict(9, 0, SPEECH_LIKE_BIAS_NORM)
    seg = make_ref.reference_segmenter(sd, 9)
    g = torch.Generator().manual_seed(3)
    clips = [torch.randn(1, 24000, generator=g), torch.randn(1, 16000, generator=g)]
    outs = seg(wav=clips, in_second=False)
    batch = torch.zeros(2, 24000)
    batch[0] = clips[0][0]
    batch[1, :16000] = clips[1][0]
    port = hubert_forward(sd, batch, [24000, 16000], 9).numpy()
    for i, o in enumerate(outs):
        assert o["hidden_states"].shape == port[i].shape
        assert np.abs(o["hidden_states"] - port[i]).max() < 2e-5           # same torch CPU kernels, different op grouping
        want = segment_ref.get_segment(o["hidden_states"], 2.6, 0.8)
        assert np.array_equal(np.asarray(o["segments"]), np.asarray(want))  # the port's get_segment == the reference's
        c = segment_ref.c_get_segment(o["hidden_states"], 2.6, 0.8)
        assert len(c) == len(want) and (len(c) == 0 or np.array_equal(c, want))
