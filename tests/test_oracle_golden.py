"""The oracle (oracle/) against outputs of the UNMODIFIED reference captured in tests/golden/ by
tools/make_golden.py.  Bit-exact everywhere: same torch CPU ops, same order."""
import numpy as np
import torch

import oracle
from conftest import split_rows


def test_sample_image_end_to_end(golden_sample, states):
    sd, sr = states
    out = oracle.infer_image(sd, sr, golden_sample["bgr"])
    assert out.dtype == np.float64 and np.array_equal(out, golden_sample["out_refined"])
    raw = oracle.infer_image(sd, None, golden_sample["bgr"])
    assert raw.dtype == np.int64 and np.array_equal(raw, golden_sample["out_raw"])
    # the known-answer vector quoted in SURVEY.md 8c / BASELINE.md
    assert out[0].tolist() == [171.375, 83.875, 1.0] and out[-1].tolist() == [163.875, 114.375, 14.0]


def test_sample_image_stages(golden_sample, states):
    sd, sr = states
    g = golden_sample
    res, st = oracle.pipeline.infer_gray(sd, sr, g["gray"], return_stages=True)
    assert np.array_equal(st["loc"][0], g["st_loc"]) and np.array_equal(st["ids"][0], g["st_ids"])
    assert np.array_equal(st["kpts"], g["st_kpts"]) and np.array_equal(st["ids_found"], g["st_ids_found"])
    assert np.array_equal(st["patches"], g["st_patches"])
    assert np.array_equal(st["heat"], g["st_heat"])
    assert np.array_equal(st["corners"], g["st_corners"]) and np.array_equal(st["refined"], g["st_refined"])


def test_decode_from_reference_logits(golden_synth):
    g = golden_synth
    counts = g["counts"]
    kp = split_rows(g["kpts"], counts)
    idf = split_rows(g["ids_found"], counts)
    for i in range(g["loc"].shape[0]):
        k, ids = oracle.pred_to_keypoints(g["loc"][i:i + 1], g["ids"][i:i + 1], 16)
        assert np.array_equal(k, kp[i]) and np.array_equal(ids, idf[i])


def test_patches_and_argmax_from_reference(golden_synth):
    g = golden_synth
    n = g["loc"].shape[0]
    counts = g["counts"][:n]
    kp = split_rows(g["kpts"], g["counts"])[:n]
    patches = split_rows(g["patches"], counts)
    heat = split_rows(g["heat"], counts)
    corners = split_rows(g["corners"], g["counts"])[:n]
    for i in range(n):
        img = oracle.pre_bgr_image(g["frames"][i])
        assert np.array_equal(oracle.extract_patches(img, kp[i]), patches[i])
        assert np.array_equal(oracle.bargmax2d(heat[i]), corners[i])


def test_synthetic_frames_end_to_end(golden_synth, states):
    sd, sr = states
    g = golden_synth
    want = split_rows(g["out_refined"], g["counts"])
    want_raw = split_rows(g["out_raw"], g["counts"])
    for i in (0, 4, 6, 15):
        assert np.array_equal(oracle.pipeline.infer_gray(sd, sr, g["frames"][i]), want[i])
        assert np.array_equal(oracle.pipeline.infer_gray(sd, None, g["frames"][i]), want_raw[i])


def test_edge_cases(golden_edge, states):
    sd, sr = states
    g = golden_edge
    want = split_rows(g["out_refined"], g["counts"])
    for i, name in enumerate(g["names"].tolist()):
        out = oracle.pipeline.infer_gray(sd, sr, g["frames"][i])
        if g["counts"][i] == 0:
            assert out.shape == (0,), name                 # np.array([]) -- inference.py:51-52
        else:
            assert np.array_equal(out, want[i]), name
    assert dict(zip(g["names"].tolist(), g["counts"].tolist()))["crowded"] == 192   # K is not capped at n_ids


def test_640x480(golden_640, states):
    sd, sr = states
    g = golden_640
    want = split_rows(g["out_refined"], g["counts"])
    assert np.array_equal(oracle.pipeline.infer_gray(sd, sr, g["frames"][0]), want[0])


def test_solve_pnp_known_answer(golden_sample):
    g = golden_sample
    ret, rvec, tvec = oracle.solve_pnp(g["out_refined"], 5, 5, 0.01, g["pnp_camera"], np.zeros(5))
    assert bool(ret) == bool(g["pnp_ret"])
    assert np.allclose(rvec, g["pnp_rvec"], atol=1e-9) and np.allclose(tvec, g["pnp_tvec"], atol=1e-9)
    assert oracle.solve_pnp(g["out_refined"][:3], 5, 5, 0.01, g["pnp_camera"], np.zeros(5)) == (False, None, None)


def test_argmax_first_max_tie_break():
    loc = np.zeros((1, 65, 1, 2), np.float32)
    ids = np.zeros((1, 17, 1, 2), np.float32)
    loc[0, 5, 0, 0] = loc[0, 9, 0, 0] = 3.0        # tie -> lowest index 5
    ids[0, 2, 0, 0] = ids[0, 7, 0, 0] = 1.0        # tie -> 2
    loc[0, 64, 0, 1] = 1.0                         # dustbin cell
    k, i = oracle.pred_to_keypoints(loc, ids, 16)
    assert k.tolist() == [[5, 0]] and i.tolist() == [2]
    heat = np.zeros((1, 64, 64), np.float32)
    heat[0, 3, 10] = heat[0, 40, 1] = 2.0
    assert oracle.bargmax2d(heat).tolist() == [[10, 3]]
