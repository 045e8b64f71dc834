"""CUDA engine vs. the committed reference outputs (tests/golden), through the public Python surface,
which calls the C ABI (dcu_infer_batch_host).  Integer results bit-exact; refined corners within 1e-3 px."""
import numpy as np
import pytest

import deepcharuco_b200 as dc
from conftest import split_rows

pytestmark = pytest.mark.gpu


def _check_refined(got, want, tol=1e-3):
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got[:, 2], want[:, 2])                       # ids bit-exact
    assert np.abs(got[:, :2] - want[:, :2]).max() <= tol               # sub-pixel within 1e-3 px


def test_sample_image_known_answer(models, golden_sample):
    deepc, refinenet = models
    g = golden_sample
    kp, img = dc.infer_image(g["bgr"], 16, deepc, refinenet)
    assert img is g["bgr"] or np.array_equal(img, g["bgr"])            # draw_pred=False returns the input image
    assert kp.dtype == np.float64
    _check_refined(kp, g["out_refined"])
    raw, _ = dc.infer_image(g["bgr"], 16, deepc, None)
    assert raw.dtype == np.int64 and np.array_equal(raw, g["out_raw"])   # integer pixels + ids: bit-exact


def test_sample_image_pnp(models, golden_sample):
    deepc, refinenet = models
    g = golden_sample
    kp, _ = dc.infer_image(g["bgr"], 16, deepc, refinenet)
    ret, rvec, tvec = dc.solve_pnp(kp, 5, 5, 0.01, g["pnp_camera"], np.zeros(5))
    assert ret and np.allclose(rvec, g["pnp_rvec"], atol=1e-6) and np.allclose(tvec, g["pnp_tvec"], atol=1e-6)
    assert dc.solve_pnp(kp[:3], 5, 5, 0.01, g["pnp_camera"], np.zeros(5)) == (False, None, None)


def test_synthetic_batch_matches_reference(models, golden_synth):
    deepc, refinenet = models
    g = golden_synth
    res = dc.infer_batch(g["frames"], 16, deepc, refinenet)
    want = split_rows(g["out_refined"], g["counts"])
    assert [0 if r.size == 0 else r.shape[0] for r in res] == g["counts"].tolist()
    for got, w in zip(res, want):
        _check_refined(got, w)
    raw = dc.infer_batch(g["frames"], 16, deepc, None)
    for got, w in zip(raw, split_rows(g["out_raw"], g["counts"])):
        assert got.dtype == np.int64 and np.array_equal(got, w)


def test_fp32_cuda_core_path_matches_reference_too(models_ffma, golden_sample, golden_synth):
    deepc, refinenet = models_ffma
    kp, _ = dc.infer_image(golden_sample["bgr"], 16, deepc, refinenet)
    _check_refined(kp, golden_sample["out_refined"])
    g = golden_synth
    for got, w in zip(dc.infer_batch(g["frames"], 16, deepc, refinenet), split_rows(g["out_refined"], g["counts"])):
        _check_refined(got, w)
    for got, w in zip(dc.infer_batch(g["frames"], 16, deepc, None), split_rows(g["out_raw"], g["counts"])):
        assert np.array_equal(got, w)


def test_batch_equals_single_frame_calls(models, golden_synth):
    deepc, refinenet = models
    frames = golden_synth["frames"][:5]
    batch = dc.infer_batch(frames, 16, deepc, refinenet)
    for f, b in zip(frames, batch):
        one = dc.infer_batch(f[None], 16, deepc, refinenet)[0]
        assert np.array_equal(one, b)


def test_edge_cases(models, golden_edge):
    deepc, refinenet = models
    g = golden_edge
    res = dc.infer_batch(g["frames"], 16, deepc, refinenet)
    want = split_rows(g["out_refined"], g["counts"])
    raw = dc.infer_batch(g["frames"], 16, deepc, None)
    want_raw = split_rows(g["out_raw"], g["counts"])
    for name, got, w, gr, wr in zip(g["names"].tolist(), res, want, raw, want_raw):
        if len(w) == 0:
            assert got.shape == (0,) and gr.shape == (0,), name       # np.array([]) for K == 0
        else:
            assert np.array_equal(gr, wr), name
            _check_refined(got, w)


def test_crowded_frame_exceeding_default_capacity(models, golden_edge):
    """192 corners in one frame > a small max_patches: the wrapper grows the workspace instead of truncating."""
    deepc, refinenet = models
    g = golden_edge
    i = g["names"].tolist().index("crowded")
    ctx = deepc._ctx
    ctx.close()
    ctx.engine(240, 320, max_batch=1, max_patches=64)
    got = dc.infer_batch(g["frames"][i:i + 1], 16, deepc, refinenet)[0]
    assert got.shape == (192, 3)
    ctx.close()


def test_640x480(models, golden_640):
    deepc, refinenet = models
    g = golden_640
    res = dc.infer_batch(g["frames"], 16, deepc, refinenet)
    for got, w in zip(res, split_rows(g["out_refined"], g["counts"])):
        _check_refined(got, w)


def test_bgr_batch_equals_gray_batch(models, golden_sample, golden_synth):
    """(N,H,W,3) BGR frames: colour conversion on the device must equal cv2.cvtColor bit for bit, hence identical results."""
    import cv2
    import torch
    from deepcharuco_b200 import _native as N
    deepc, refinenet = models
    rng = np.random.default_rng(3)
    gray = golden_synth["frames"][:4]
    bgr = np.stack([cv2.cvtColor(f, cv2.COLOR_GRAY2BGR) for f in gray])
    bgr = np.clip(bgr.astype(np.int16) + rng.integers(-20, 21, bgr.shape), 0, 255).astype(np.uint8)     # real colour content
    bgr = np.concatenate([bgr, golden_sample["bgr"][None]], 0)
    want_gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in bgr])
    eng = deepc._ctx.engine(240, 320, max_batch=8)
    d_bgr = torch.from_numpy(bgr).cuda()
    d_gray = torch.empty((5, 240, 320), dtype=torch.uint8, device="cuda")
    N.check(N.lib().dcu_bgr_to_gray(eng.handle, d_bgr.data_ptr(), 5, d_gray.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(d_gray.cpu().numpy(), want_gray)
    a = dc.infer_batch(bgr, 16, deepc, refinenet)
    b = dc.infer_batch(want_gray, 16, deepc, refinenet)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    _check_refined(a[4], golden_sample["out_refined"])


def test_empty_batch_and_bad_shapes(models):
    deepc, refinenet = models
    assert dc.infer_batch(np.zeros((0, 240, 320), np.uint8), 16, deepc, refinenet) == []
    with pytest.raises(Exception):
        dc.infer_batch(np.zeros((1, 100, 101), np.uint8), 16, deepc, refinenet)     # not a multiple of 8


def test_device_resize_is_bit_exact_with_cv2_and_feeds_the_pipeline(models, golden_sample):
    """SURVEY 8f row 1, the optional resize: cv2.resize(frame, (320, 240), INTER_LINEAR) (inference.py:131-132) on the device, bit-exact
    with cv2 when shrinking; infer_batch(..., input_size) = resize + BGR->gray + pipeline equals the host-resized call."""
    import cv2
    from deepcharuco_b200.inference import resize_gpu
    from deepcharuco_b200 import _native as N
    deepc, refinenet = models
    rng = np.random.default_rng(3)
    for (Hs, Ws, ch) in ((1920, 2560, 3), (480, 640, 3), (720, 1280, 1), (300, 421, 3), (241, 323, 1), (240, 320, 3)):
        src = rng.integers(0, 256, (2, Hs, Ws, ch) if ch == 3 else (2, Hs, Ws)).astype(np.uint8)
        got = resize_gpu(src, (320, 240))
        for i in range(2):
            assert np.array_equal(got[i], cv2.resize(src[i], (320, 240), interpolation=cv2.INTER_LINEAR)), (Hs, Ws, ch)
    with pytest.raises(N.DcuError):
        resize_gpu(np.zeros((1, 120, 160), np.uint8), (320, 240))            # enlarging is not bit-exact with cv2: refused
    big = cv2.resize(golden_sample["bgr"], (1280, 960), interpolation=cv2.INTER_CUBIC)      # a "camera frame" of the sample scene
    small = cv2.resize(big, (320, 240), interpolation=cv2.INTER_LINEAR)
    a = dc.infer_batch(big[None], 16, deepc, refinenet, input_size=(320, 240))[0]
    b = dc.infer_batch(small[None], 16, deepc, refinenet)[0]
    assert a.shape == b.shape and a.shape[0] >= 6 and np.array_equal(a, b)
