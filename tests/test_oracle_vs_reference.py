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


def test_metrics_oracles_match_live_reference_classes(ref_models):
    """oracle/metrics.py against the reference's own DC_Metrics / Refinenet_Metrics on fresh random inputs (beyond the committed goldens):
    random arg-max maps with a handful of labelled cells, unique label ids (the reference raises on a repeated label id)."""
    rh.load_reference()
    from models import metrics as M
    rng = np.random.default_rng(5)
    n, h, w = 6, 30, 40
    loc_a = rng.integers(0, 65, size=(n, h, w)); ids_a = np.full((n, h, w), 16)
    loc_t = np.full((n, h, w), 64); ids_t = np.full((n, h, w), 16)
    for i in range(n):
        cells = rng.choice(h * w, size=24, replace=False)
        for j, c in enumerate(cells[:14]):                    # predictions: ids may repeat
            ids_a[i].flat[c] = int(rng.integers(0, 16)); loc_a[i].flat[c] = int(rng.integers(0, 64))
        for j, c in enumerate(cells[8:8 + int(rng.integers(0, 12))]):     # labels: unique ids, partly on predicted cells
            ids_t[i].flat[c] = j; loc_t[i].flat[c] = int(rng.integers(0, 64))
    one_hot = lambda a, c: (np.arange(c)[None, :, None, None] == a[:, None]).astype(np.float32)
    loc_hat, ids_hat = one_hot(loc_a, 65), one_hot(ids_a, 17)
    ref_m, ora_m = M.DC_Metrics(16), oracle.metrics.DCMetrics(16)
    ref_m.update((torch.from_numpy(loc_hat), torch.from_numpy(ids_hat)), (torch.from_numpy(loc_t), torch.from_numpy(ids_t)))
    ora_m.update((loc_hat, ids_hat), (loc_t, ids_t))
    d, r = ref_m.compute()
    assert np.allclose(ora_m.compute(), [float(d), float(r)], rtol=1e-6, atol=1e-6)
    heat = rng.standard_normal((9, 64, 64)).astype(np.float32)
    tgt = rng.standard_normal((9, 64, 64)).astype(np.float32)
    ref_r, ora_r = M.Refinenet_Metrics(), oracle.metrics.RefinenetMetrics()
    ref_r.update(torch.from_numpy(heat[:, None]), torch.from_numpy(tgt))
    ora_r.update(heat[:, None], tgt)
    assert np.allclose(ora_r.compute(), float(ref_r.compute()), rtol=1e-6)
