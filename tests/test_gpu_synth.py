"""Frames generated on the device (csrc/synth.cu): the warp against cv2 itself, whole frames against the
numpy mirror (oracle/synth.py) bit for bit, and the statistics that make the frames useful (boards the detector finds, at the
ground-truth positions)."""
import numpy as np
import pytest

import deepcharuco_b200 as dc
from deepcharuco_b200 import synth
from oracle import synth as S
from test_oracle_synth import warp_cases

pytestmark = pytest.mark.gpu


def test_device_warp_is_bit_exact_with_cv2():
    import cv2
    for src, M, dsize in warp_cases():
        want = cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_LINEAR)
        got = synth.warp_perspective_u8_gpu(src, M, dsize)
        assert np.array_equal(want, got), (dsize, int((want != got).sum()))


def _frame_dict(p, i):
    return dict(lat_step=int(p["lat_step"][i]), bg_lo=p["bg_lo"][i], bg_hi=p["bg_hi"][i], gain=p["gain"][i], blur_w=p["blur_w"][i],
                Hinv=list(p["Hinv"][i]), H=list(p["H"][i]), corners=list(p["corners"][i]))


@pytest.mark.parametrize("H,W,nb,seed", [(240, 320, 1, 5), (480, 640, 4, 9), (200, 296, 2, 1)])
def test_frames_are_bit_identical_with_the_numpy_mirror(H, W, nb, seed):
    n = 5
    frames, corners = synth.make_frames_gpu(n, H, W, seed=seed, n_boards=nb, first_index=2)
    p = synth.gpu_frame_params(n, H, W, seed, nb, first_index=2)
    board = synth.board_render(240)
    for i in range(n):
        want, wc = S.make_frame(board, seed, 2 + i, H, W, nb, params=_frame_dict(p, i))
        assert np.array_equal(frames[i], want), (i, int((frames[i] != want).sum()))
        own, _ = S.make_frame(board, seed, 2 + i, H, W, nb)                    # the oracle's own parameters: same frame up to rare
        assert (own != want).mean() < 1e-3                                      # fixed-point rounding flips from 1e-13 parameter noise
    assert np.array_equal(corners, p["corners"])


def test_frames_depend_only_on_seed_and_index():
    a, _ = synth.make_frames_gpu(12, seed=3)
    b, _ = synth.make_frames_gpu(4, seed=3, first_index=5)
    assert np.array_equal(a[5:9], b)
    c, _ = synth.make_frames_gpu(4, seed=4, first_index=5)
    assert not np.array_equal(b, c)
    d = synth.make_frames_gpu(4, seed=3, first_index=5, return_device=True)[0]
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), b)


def test_generated_frames_feed_the_engine(models):
    """64 generated frames through the engine: the boards are found (K like the host generator's, SURVEY.md 8d: mean ~14, max 17)
    and the refined corners sit on the ground-truth corners (pixel_error against the generator's labels, utils.py:33-52)."""
    from deepcharuco_b200.metrics import pixel_error_batch
    deepc, refinenet = models
    frames, corners = synth.make_frames_gpu(64, seed=21)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    ks = np.array([0 if r.size == 0 else r.shape[0] for r in refined])
    assert ks.mean() > 11 and ks.max() <= 17, ks
    labels = synth.corner_labels(corners, 240, 320)
    status, out = pixel_error_batch(raw, refined, labels)
    ok = status == 1
    assert ok.sum() >= 48
    assert out[ok, 1].mean() < 0.8 and out[ok, 0].mean() < 1.0           # mean refined / raw error in pixels against the ground truth
    assert 0.3 < frames.std() / 64.0 < 1.2 and 30 < frames.mean() < 160
