"""Live cross-check of the oracle against the mounted reference (this container only; skipped on the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import ref_harness as rh  # noqa: E402
import oracle  # noqa: E402
from deepcharuco_b200 import synth  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref_models():
    return rh.load_reference(), rh.load_reference_models("cpu")


def test_logits_bit_identical(ref_models, states):
    ref, (deepc, refinenet) = ref_models
    sd, sr = states
    frames = synth.make_frames(2, seed=11)
    for f in frames:
        x = torch.from_numpy(oracle.pre_bgr_image(f))
        loc_r, ids_r = deepc.infer_image(x)
        loc_o, ids_o = oracle.detector_forward(sd, x[None])
        assert torch.equal(loc_r, loc_o) and torch.equal(ids_r, ids_o)


def test_end_to_end_identical(ref_models, states):
    import cv2
    ref, (deepc, refinenet) = ref_models
    sd, sr = states
    for f in synth.make_frames(3, seed=12):
        bgr = cv2.cvtColor(f, cv2.COLOR_GRAY2BGR)
        a, _ = ref.infer_image(bgr, 16, deepc, refinenet)
        b = oracle.infer_image(sd, sr, bgr)
        assert a.dtype == b.dtype and np.array_equal(a, b)


def test_converted_weights_match_checkpoints(states):
    from deepcharuco_b200 import weights_io as W
    for ckpt, st in ((rh.DEEPC_CKPT, states[0]), (rh.REFINE_CKPT, states[1])):
        live = W.load_state(ckpt)
        assert live.keys() == st.keys()
        assert all(np.array_equal(live[k], st[k]) for k in live)
