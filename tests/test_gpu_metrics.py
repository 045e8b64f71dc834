"""DC_Metrics on the device against the reference's own DC_Metrics outputs
(tests/golden/metrics_seed0.npz) and the oracle restatement, through dcu_decode_gather + dcu_dc_metrics."""
import numpy as np
import pytest
import torch

import deepcharuco_b200 as dc
import oracle
from conftest import load_golden
from deepcharuco_b200.metrics import DC_Metrics, Refinenet_Metrics

pytestmark = pytest.mark.gpu
TOL = 1e-6      # fp32: per-id distances are bit-exact (integer coordinates); only the order of the final sums differs


@pytest.fixture(scope="module")
def models():
    return dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda")


def _one_hot(arg, c):
    return torch.from_numpy((np.arange(c)[None, :, None, None] == arg[:, None]).astype(np.float32)).cuda()


def test_update_on_logits_matches_reference(models):
    g = load_golden("metrics_seed0.npz")
    m = DC_Metrics(16, models[0])
    loc_hat, ids_hat = _one_hot(g["loc_argmax"], 65), _one_hot(g["ids_argmax"], 17)
    l2, ratio, valid = m.update((loc_hat, ids_hat), (g["loc_target"], g["ids_target"]))
    want_valid = ~np.isnan(g["per_l2"])
    assert np.array_equal(valid, want_valid)
    assert np.allclose(l2[valid], g["per_l2"][want_valid], rtol=TOL, atol=TOL) and np.allclose(ratio[valid], g["per_ratio"][want_valid], rtol=TOL, atol=TOL)
    assert np.allclose(m.compute(), g["after_update1"], rtol=TOL, atol=TOL)
    m.update((loc_hat[3:8], ids_hat[3:8]), (g["loc_target"][3:8], g["ids_target"][3:8]))
    assert np.allclose(m.compute(), g["after_update2"], rtol=TOL, atol=TOL)


def test_update_frames_end_to_end(models, golden_synth):
    """Frames -> the engine's detector + decode -> metric: equals the reference's metric on the reference's logits, because the
    engine's kept cells / ids / pixels are bit-exact."""
    g = load_golden("metrics_seed0.npz")
    m = DC_Metrics(16, models[0])
    l2, ratio, valid = m.update_frames(golden_synth["frames"], (g["loc_target"], g["ids_target"]))
    want_valid = ~np.isnan(g["per_l2"])
    assert np.array_equal(valid, want_valid)
    assert np.allclose(l2[valid], g["per_l2"][want_valid], rtol=TOL, atol=TOL)
    assert np.allclose(ratio[valid], g["per_ratio"][want_valid], rtol=TOL, atol=TOL)
    assert np.allclose(m.compute(), g["after_update1"], rtol=TOL, atol=TOL)
    # and the oracle on the same inputs
    o = oracle.metrics.DCMetrics(16)
    o.update((_one_hot(g["loc_argmax"], 65).cpu().numpy(), _one_hot(g["ids_argmax"], 17).cpu().numpy()), (g["loc_target"], g["ids_target"]))
    assert np.allclose(m.compute(), o.compute(), rtol=TOL, atol=TOL)


def test_refinenet_metrics_match_reference(models, golden_synth):
    """Refinenet_Metrics on the device against the reference's own class (tests/golden/refinenet_metrics_seed0.npz): heat-map inputs
    like the reference's update(), and the engine's own RefineNet on the golden patches (same arg-maxes as the reference's)."""
    r = load_golden("refinenet_metrics_seed0.npz")
    heat = golden_synth["heat"]
    p = heat.shape[0]
    target = np.stack([np.roll(heat[i], (int(r["shifts"][i, 0]), int(r["shifts"][i, 1])), axis=(0, 1)) for i in range(p)])
    m = Refinenet_Metrics(models[1])
    d = m.update(torch.from_numpy(heat[:, None]).cuda(), target)
    assert np.array_equal(d, r["per_dist"])                       # integer arg-max positions: the fp32 distances are bit-exact
    assert np.allclose(m.compute(), r["after_update1"], rtol=TOL)
    m.update(heat[5:20, None], target[5:20])
    assert np.allclose(m.compute(), r["after_update2"], rtol=TOL)
    # first-maximum tie rule on a flat map, and a maximum in the last element
    flat = np.zeros((2, 64, 64), np.float32); flat[1, 63, 63] = 1.0
    tgt = np.zeros((2, 64, 64), np.float32); tgt[0, 3, 4] = 1.0; tgt[1, 63, 63] = 2.0
    d2 = Refinenet_Metrics(models[1]).update(flat, tgt)
    assert np.array_equal(d2, np.array([5.0, 0.0], np.float32))
    m2 = Refinenet_Metrics(models[1])
    d3 = m2.update_patches(golden_synth["patches"], golden_synth["kpts"][:p], target)
    assert np.array_equal(d3, r["per_dist"])
    want = oracle.metrics.RefinenetMetrics.per_sample(heat, target)
    assert np.array_equal(d3, want)


def test_dense_predictions_do_not_overrun_capacity(models):
    """Random logits keep ~1100 of 1200 cells per frame (an untrained detector; the reference's own metrics.py self-test feeds
    exactly this): far more than the default 64 corners per frame of an inference engine.  `update` must size the decode for every
    cell and agree with the oracle restatement of DC_Metrics; the kernels must never read past the row buffer."""
    rng = np.random.default_rng(11)
    n = 6
    loc = rng.standard_normal((n, 65, 30, 40)).astype(np.float32)
    ids = rng.standard_normal((n, 17, 30, 40)).astype(np.float32)
    loc_t = np.full((n, 30, 40), 64, np.int64)
    ids_t = np.full((n, 30, 40), 16, np.int64)
    for f in range(n):
        cells = rng.choice(1200, 16, replace=False)
        for i, c in enumerate(cells):
            ids_t[f, c // 40, c % 40] = i
            loc_t[f, c // 40, c % 40] = rng.integers(0, 64)
    m = DC_Metrics(16, models[0])
    l2, ratio, valid = m.update((torch.from_numpy(loc).cuda(), torch.from_numpy(ids).cuda()), (loc_t, ids_t))
    o = oracle.metrics.DCMetrics(16)
    o.update((loc, ids), (loc_t, ids_t))
    assert valid.all()
    assert np.allclose(m.compute(), o.compute(), rtol=1e-5, atol=1e-5)
    kp, _ = oracle.pred_to_keypoints(loc, ids, 16)
    assert kp.shape[0] > 64 * n * 10          # the case really is dense


def test_capacity_is_an_error_not_an_overrun(states):
    """An engine with room for 8 corners, fed frames with ~15 each: the device entry point reports DCU_ERR_CAPACITY (RefineNet leg)
    and the metric / pose kernels stay inside the 8 rows."""
    from deepcharuco_b200 import _native as N, synth
    frames = synth.make_frames(4, 240, 320, seed=1)
    e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=4, max_patches=8)
    try:
        fr = torch.from_numpy(frames).cuda()
        with pytest.raises(N.CapacityError):
            e.infer_batch_device(fr.data_ptr(), 4, 16, True, None)
        o = e.infer_batch_device(fr.data_ptr(), 4, 16, False, None)
        torch.cuda.synchronize()
        assert int(o["total"].item()) > 8
        ret, rvec, tvec = e.solve_pnp_batch_device(4, 5, 5, 0.01, np.array([[300., 0, 160], [0, 300., 120], [0, 0, 1]]), None, use_refined=False)
        torch.cuda.synchronize()
        assert ret.shape[0] == 4 and bool(torch.isfinite(rvec[ret.bool()]).all())
        with pytest.raises(N.CapacityError):
            e.infer_batch_host(frames, 16, True)
    finally:
        e.close()


def test_pixel_error_batch_is_bit_identical_with_the_reference(models):
    """utils.pixel_error (utils.py:33-52) for a whole batch in one launch: float64 means / maxima bit-identical with the numbers the
    unmodified reference produced (tests/golden/pixel_error_seed0.npz), same skip rules; single-frame drop-in returns the same pair."""
    from test_oracle_metrics import pixel_error_case
    from deepcharuco_b200.metrics import pixel_error, pixel_error_batch
    raws, refs, targets, p = pixel_error_case()
    status, out = pixel_error_batch(raws, refs, targets)
    assert np.array_equal(status, p["status"].astype(np.int32))
    ok = status == 1
    assert ok.sum() >= 10 and np.array_equal(out[ok], p["out"][ok])
    i = int(np.nonzero(ok)[0][0])
    a, b = pixel_error(raws[i], refs[i], targets[i], verbose=False)
    assert (a, b) == (p["out"][i, 0], p["out"][i, 1])
    j = int(np.nonzero(p["status"] == 0)[0][0])
    assert pixel_error(raws[j], refs[j], targets[j], verbose=False) == (None, None) or raws[j].shape[0] == 0
    # several predictions AND several labels of one id in different numbers: numpy raises in the reference
    raw = np.array([[10, 10, 1], [20, 20, 1], [30, 30, 1]], np.int64)
    tgt = np.array([[10.0, 10.0, 1.0], [21.0, 20.0, 1.0]])
    st, _ = pixel_error_batch([raw], [raw.astype(np.float64)], [tgt])
    assert st[0] == -1


def test_pixel_error_on_engine_results(models, golden_synth):
    """End to end: frames -> engine (raw + refined) -> pixel error against labels = the reference's own refined corners: the
    refined error is 0 wherever the engine's arg-max equals the reference's, the raw error is the sub-pixel offset (< 0.75 px)."""
    from deepcharuco_b200.metrics import pixel_error_batch
    deepc, refinenet = models
    g = golden_synth
    from conftest import split_rows
    refined = dc.infer_batch(g["frames"], 16, deepc, refinenet)
    raw = dc.infer_batch(g["frames"], 16, deepc, None)
    labels = split_rows(g["out_refined"], g["counts"])
    status, out = pixel_error_batch(raw, refined, labels)
    assert (status == 1).all()
    assert out[:, 1].max() <= 1e-3 and 0.2 < out[:, 0].mean() < 0.75
    assert np.array_equal(out[:, 0], out[:, 2])          # refined == label, so raw-vs-label == refined-vs-raw (symmetric distances)
