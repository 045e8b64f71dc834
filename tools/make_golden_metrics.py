"""Generate tests/golden/metrics_seed0.npz by running the UNMODIFIED reference's validation metric
(/root/reference/src/models/metrics.py: DC_Metrics, label_to_keypoints, pred_to_keypoints) in this container.

    python tools/make_golden_metrics.py

Inputs: the reference detector's logits on the 16 golden frames (tests/golden/synthetic_320x240_seed0.npz) and label maps
fabricated from the reference's own decode of those logits (some corners moved inside their cell, some dropped, one spurious
id added, one frame emptied), so that every branch of the matching logic is exercised.  Stored: the label maps, the
reference's per-sample distance / ratio and the accumulated DC_Metrics state after two update() calls.  The logits themselves
are NOT stored (the GPU test feeds the frames to the engine, whose decode is bit-exact with the reference's)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def main():
    torch.set_num_threads(8)
    rh.load_reference()
    from models import metrics as M                       # the reference's metrics.py, unmodified
    from models.model_utils import pre_bgr_image
    deepc, _ = rh.load_reference_models("cpu")
    g = np.load(os.path.join(OUT, "synthetic_320x240_seed0.npz"))
    frames = g["frames"]
    n = frames.shape[0]
    locs, idss = [], []
    for f in frames:
        loc, ids = deepc.infer_image(torch.tensor(pre_bgr_image(f)))
        locs.append(loc[0]); idss.append(ids[0])
    loc_hat, ids_hat = torch.stack(locs), torch.stack(idss)
    loc_arg, ids_arg = M.pred_argmax(loc_hat, ids_hat, 16)
    rng = np.random.default_rng(0)
    loc_t, ids_t = loc_arg.clone(), ids_arg.clone()
    loc_t[ids_t == 16] = 64
    for i in range(n):
        cells = torch.argwhere(ids_t[i] != 16)
        for (r, c) in cells.tolist():
            u = rng.random()
            if u < 0.15:
                ids_t[i, r, c] = 16; loc_t[i, r, c] = 64                     # missed by the labels
            elif u < 0.55:
                loc_t[i, r, c] = int(rng.integers(0, 64))                    # label elsewhere in the cell (<= ~10 px away)
        empty = torch.argwhere(ids_t[i] == 16)
        r, c = empty[int(rng.integers(0, len(empty)))].tolist()
        free = sorted(set(range(16)) - set(ids_t[i][ids_t[i] != 16].tolist()))   # label ids are unique per sample (the reference's
        if free:                                                                 # metric raises on a repeated target id)
            ids_t[i, r, c] = int(free[int(rng.integers(0, len(free)))]); loc_t[i, r, c] = int(rng.integers(0, 64))
    ids_t[5] = 16; loc_t[5] = 64                                             # a sample without labels: distance is None
    met = M.DC_Metrics(16)
    per_l2, per_ratio = [], []
    for i in range(n):
        kp, idd = M.pred_to_keypoints(loc_hat[i:i + 1], ids_hat[i:i + 1], 16)
        kt, it = M.label_to_keypoints(loc_t[i:i + 1], ids_t[i:i + 1], 16)
        l2 = met.compute_l2_distance(kp, idd, kt, it)
        ra = met.compute_ratio(kp, idd, kt, it)
        per_l2.append(np.nan if l2 is None else float(l2)); per_ratio.append(np.nan if ra is None else float(ra))
    met.update((loc_hat, ids_hat), (loc_t, ids_t))
    d1, r1 = float(met.distance), float(met.ratio)
    met.update((loc_hat[3:8], ids_hat[3:8]), (loc_t[3:8], ids_t[3:8]))
    d2, r2 = met.compute()
    np.savez_compressed(os.path.join(OUT, "metrics_seed0.npz"), loc_target=loc_t.numpy().astype(np.int64),
                        ids_target=ids_t.numpy().astype(np.int64), loc_argmax=loc_arg.numpy().astype(np.int64),
                        ids_argmax=ids_arg.numpy().astype(np.int64), per_l2=np.array(per_l2, np.float32),
                        per_ratio=np.array(per_ratio, np.float32), after_update1=np.array([d1, r1], np.float32),
                        after_update2=np.array([float(d2), float(r2)], np.float32),
                        meta=np.array("reference metrics.py DC_Metrics(16), torch " + torch.__version__))
    print("per-sample l2:", np.round(per_l2, 3)); print("per-sample ratio:", np.round(per_ratio, 3)); print(d1, r1, float(d2), float(r2))

    # --- Refinenet_Metrics (metrics.py:135-161): predictions = the reference RefineNet's heat maps of the golden patches, targets =
    # the same maps rolled by a few pixels (stored as the shifts only; tests rebuild them), so arg-max positions differ by known amounts
    heat = torch.from_numpy(g["heat"])                                  # (P, 64, 64)
    P = heat.shape[0]
    shifts = rng.integers(-6, 7, size=(P, 2))
    shifts[::5] = 0                                                     # some exact hits
    target = torch.stack([torch.roll(heat[i], (int(shifts[i, 0]), int(shifts[i, 1])), dims=(0, 1)) for i in range(P)])
    rm = M.Refinenet_Metrics()
    per = []
    for i in range(P):
        one = M.Refinenet_Metrics()
        one.update(heat[i:i + 1, None], target[i:i + 1])
        per.append(float(one.compute()))
    rm.update(heat[:, None], target)
    u1 = float(rm.compute())
    rm.update(heat[5:20, None], target[5:20])
    u2 = float(rm.compute())
    np.savez_compressed(os.path.join(OUT, "refinenet_metrics_seed0.npz"), shifts=shifts.astype(np.int64), per_dist=np.array(per, np.float32),
                        after_update1=np.float32(u1), after_update2=np.float32(u2),
                        meta=np.array("reference metrics.py Refinenet_Metrics on tests/golden/synthetic_320x240_seed0.npz heat maps, torch " + torch.__version__))
    print("refinenet metric:", u1, u2, np.round(per[:8], 3))


if __name__ == "__main__":
    main()
