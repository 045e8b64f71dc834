"""oracle/metrics.py against the outputs of the reference's own DC_Metrics (tests/golden/metrics_seed0.npz,
tools/make_golden_metrics.py).  Predictions enter as the reference's arg-max maps (one-hot logits give the same decode)."""
import numpy as np

import oracle
from conftest import load_golden


def _one_hot(arg, c):
    return (np.arange(c)[None, :, None, None] == arg[:, None]).astype(np.float32)


def test_metrics_oracle_matches_reference():
    g = load_golden("metrics_seed0.npz")
    loc_hat, ids_hat = _one_hot(g["loc_argmax"], 65), _one_hot(g["ids_argmax"], 17)
    m = oracle.metrics.DCMetrics(16)
    for i in range(loc_hat.shape[0]):
        l2, ratio = m.sample(loc_hat[i], ids_hat[i], g["loc_target"][i], g["ids_target"][i])
        if np.isnan(g["per_l2"][i]):
            assert l2 is None and ratio is None            # a sample without labels (metrics.py:80-81, :106-107)
        else:
            assert abs(l2 - g["per_l2"][i]) <= 1e-6 * max(1.0, g["per_l2"][i]) and abs(ratio - g["per_ratio"][i]) <= 1e-6
    m.update((loc_hat, ids_hat), (g["loc_target"], g["ids_target"]))
    assert np.allclose(m.compute(), g["after_update1"], rtol=1e-6, atol=1e-6)
    m.update((loc_hat[3:8], ids_hat[3:8]), (g["loc_target"][3:8], g["ids_target"][3:8]))
    assert np.allclose(m.compute(), g["after_update2"], rtol=1e-6, atol=1e-6)


def _refinenet_case():
    g = load_golden("synthetic_320x240_seed0.npz")
    r = load_golden("refinenet_metrics_seed0.npz")
    heat = g["heat"]
    target = np.stack([np.roll(heat[i], (int(r["shifts"][i, 0]), int(r["shifts"][i, 1])), axis=(0, 1)) for i in range(len(heat))])
    return heat, target, r


def test_refinenet_metrics_oracle_matches_reference():
    heat, target, r = _refinenet_case()
    m = oracle.metrics.RefinenetMetrics()
    assert np.allclose(m.per_sample(heat, target), r["per_dist"], rtol=1e-6, atol=1e-6)
    m.update(heat[:, None], target)
    assert np.allclose(m.compute(), r["after_update1"], rtol=1e-6)
    m.update(heat[5:20, None], target[5:20])
    assert np.allclose(m.compute(), r["after_update2"], rtol=1e-6)


def pixel_error_case():
    """(raw, refined, target) per frame + the reference's status / numbers (tests/golden/pixel_error_seed0.npz, made by
    tools/make_golden_pixel_error.py from the unmodified reference's utils.pixel_error)."""
    from conftest import split_rows
    raws, refs = [], []
    for name in ("synthetic_320x240_seed0.npz", "edge_cases.npz"):
        g = load_golden(name)
        raws += split_rows(g["out_raw"], g["counts"]); refs += split_rows(g["out_refined"], g["counts"])
    p = load_golden("pixel_error_seed0.npz")
    targets = split_rows(p["targets"], p["target_counts"])
    return raws, refs, targets, p


def test_pixel_error_oracle_matches_reference():
    raws, refs, targets, p = pixel_error_case()
    assert len(raws) == len(targets) == len(p["status"]) and p["status"].sum() >= 10
    for raw, ref, t, st, out in zip(raws, refs, targets, p["status"], p["out"]):
        if raw.shape[0] == 0 or t.shape[0] == 0:
            assert st == 0
            continue
        got_st, got = oracle.metrics.pixel_error(raw, ref, t)
        assert got_st == st
        if st:
            assert np.array_equal(got, out)            # float64, bit-identical with the reference
