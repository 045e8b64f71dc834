"""The reference's Python surface below `infer_image`, called the way /root/reference/src/inference.py:41-60 calls it:
`deepc.infer_image` (net.py:127), `pred_to_keypoints` / `extract_patches` / `pre_bgr_image` (model_utils.py, imported by name at
inference.py:10) and `refinenet.infer_patches` (refinenet.py:143), checked stage by stage against the outputs of the unmodified
reference on its own sample image (tests/golden/sample_image.npz: st_loc, st_ids, st_kpts, st_ids_found, st_patches, st_refined,
st_corners) and against the oracle for batches and other n_ids."""
import numpy as np
import pytest
import torch

import deepcharuco_b200 as dc
import oracle
from conftest import split_rows

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.1        # |delta| on logits of magnitude ~1e2 (tests/test_gpu_stages.py)


def test_reference_call_sequence_on_the_sample_image(models, golden_sample):
    """inference.py:41-60 line by line with this package's objects instead of the reference's."""
    deepc, refinenet = models
    g = golden_sample
    img_gray = dc.pre_bgr_image(g["gray"])                                     # :41
    assert img_gray.dtype == np.float32 and img_gray.shape == (1, 240, 320)
    assert np.array_equal(img_gray, oracle.pre_bgr_image(g["gray"]))
    img_t = torch.tensor(img_gray, device="cuda")                              # :42
    loc_hat, ids_hat = deepc.infer_image(img_t)                                # :43
    assert tuple(loc_hat.shape) == (1, 65, 30, 40) and tuple(ids_hat.shape) == (1, 17, 30, 40)
    assert loc_hat.is_cuda and loc_hat.dtype == torch.float32
    assert np.abs(loc_hat.cpu().numpy()[0] - g["st_loc"]).max() < LOGIT_TOL
    assert np.abs(ids_hat.cpu().numpy()[0] - g["st_ids"]).max() < LOGIT_TOL
    kpts_hat, ids_found = dc.pred_to_keypoints(loc_hat, ids_hat, 16)           # :44
    assert kpts_hat.dtype == torch.int64 and ids_found.dtype == torch.int64
    assert np.array_equal(kpts_hat.cpu().numpy(), g["st_kpts"])                # row-major order of the reference's nonzero()
    assert np.array_equal(ids_found.cpu().numpy(), g["st_ids_found"])
    patches = dc.extract_patches(img_t, kpts_hat)                              # :55
    assert np.array_equal(patches.cpu().numpy(), g["st_patches"])              # fp32 patches bit-exact
    refined, corners = refinenet.infer_patches(patches, kpts_hat)              # :57
    assert refined.dtype == torch.float32 and corners.dtype == torch.int64
    assert np.array_equal(corners.cpu().numpy(), g["st_corners"])
    assert np.array_equal(refined.cpu().numpy(), g["st_refined"])
    keypoints = refined.cpu().numpy()                                          # :59-60, :68-70
    ids_np = ids_found.cpu().numpy()
    out = np.array([[k[0], k[1], i] for k, i in sorted(zip(keypoints, ids_np), key=lambda x: x[1])])
    assert np.abs(out - g["out_refined"]).max() <= 1e-3


def test_handles_accept_numpy_and_4d_patches(models, golden_sample):
    deepc, refinenet = models
    g = golden_sample
    loc_a, ids_a = deepc.infer_image(dc.pre_bgr_image(g["gray"]))              # ndarray in, CUDA tensors out
    loc_b, ids_b = deepc.infer_image(torch.from_numpy(dc.pre_bgr_image(g["gray"])))
    assert torch.equal(loc_a, loc_b) and torch.equal(ids_a, ids_b)
    r3, c3 = refinenet.infer_patches(g["st_patches"], g["st_kpts"])
    r4, c4 = refinenet.infer_patches(torch.from_numpy(g["st_patches"])[:, None], torch.from_numpy(g["st_kpts"]))   # (K,1,24,24)
    assert torch.equal(r3, r4) and torch.equal(c3, c4)
    assert np.array_equal(r3.cpu().numpy(), g["st_refined"])
    with pytest.raises(AssertionError):                                        # refinenet.py:102
        refinenet.infer_patches(np.zeros((2, 20, 20), np.float32), np.zeros((2, 2), np.int64))
    with pytest.raises(AssertionError):
        deepc.infer_image(np.zeros((240, 320), np.float32))


def test_pred_to_keypoints_on_a_batch_drops_the_batch_index(golden_synth):
    """model_utils.py:121-122 uses only the last two index columns: a batch decodes to ONE row-major list over (frame, cell)."""
    g = golden_synth
    loc, ids = torch.from_numpy(g["loc"]).cuda(), torch.from_numpy(g["ids"]).cuda()
    kp, idf = dc.pred_to_keypoints(loc, ids, 16)
    wk, wi = oracle.pred_to_keypoints(g["loc"], g["ids"], 16)
    assert np.array_equal(kp.cpu().numpy(), wk) and np.array_equal(idf.cpu().numpy(), wi)
    n = g["loc"].shape[0]
    assert kp.shape[0] == int(g["counts"][:n].sum())
    assert np.array_equal(kp.cpu().numpy(), np.concatenate(split_rows(g["kpts"], g["counts"])[:n]))


@pytest.mark.parametrize("n_ids", [9, 30])
def test_pred_to_keypoints_other_n_ids(n_ids):
    """Random logits (every cell a candidate, ~1/(n_ids+1) dropped by the ids dustbin, 1/65 by the loc dustbin) incl. exact ties."""
    rng = np.random.default_rng(n_ids)
    loc = rng.standard_normal((2, 65, 30, 40)).astype(np.float32)
    ids = rng.standard_normal((2, n_ids + 1, 30, 40)).astype(np.float32)
    loc[0, 7, 3, 3] = loc[0, 20, 3, 3] = 9.0            # tie -> first index
    ids[1, 2, 5, 5] = ids[1, 4, 5, 5] = 9.0
    kp, idf = dc.pred_to_keypoints(torch.from_numpy(loc).cuda(), torch.from_numpy(ids).cuda(), n_ids)
    wk, wi = oracle.pred_to_keypoints(loc, ids, n_ids)
    assert wk.shape[0] > 2000
    assert np.array_equal(kp.cpu().numpy(), wk) and np.array_equal(idf.cpu().numpy(), wi)


def test_extract_patches_borders_and_other_sizes():
    rng = np.random.default_rng(5)
    for (H, W) in ((240, 320), (480, 640), (72, 136)):
        img = rng.standard_normal((1, H, W)).astype(np.float32)
        kp = np.array([[0, 0], [W - 1, H - 1], [W // 2, H // 2], [3, H - 2], [W - 5, 7]], np.int64)
        out = dc.extract_patches(torch.from_numpy(img).cuda(), torch.from_numpy(kp).cuda())
        assert np.array_equal(out.cpu().numpy(), oracle.extract_patches(img, kp))


def test_infer_image_with_draw_pred_returns_an_annotated_copy(models, golden_sample):
    deepc, refinenet = models
    g = golden_sample
    kp, img = dc.infer_image(g["bgr"], 16, deepc, refinenet, draw_pred=True)
    assert img is not g["bgr"] and img.shape == g["bgr"].shape and not np.array_equal(img, g["bgr"])
    assert np.abs(kp - g["out_refined"]).max() <= 1e-3
