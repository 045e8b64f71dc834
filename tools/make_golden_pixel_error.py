"""Generate tests/golden/pixel_error_seed0.npz by running the UNMODIFIED reference's `utils.pixel_error` /
`utils.compute_l2_distance` (/root/reference/src/utils.py:6-52) in this container.

    python tools/make_golden_pixel_error.py

Inputs: the reference's own raw / refined keypoints of the golden frames (tests/golden/synthetic_320x240_seed0.npz and
edge_cases.npz, incl. the 12-board frame with 12 predictions per id) and fabricated float labels: the refined corner moved by a
random sub-pixel offset, some labels dropped (-> the reference returns (None, None)), some extra labels nobody predicted (their
distance stays 0), one frame without labels.  Stored: the labels and the reference's numbers."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def split_rows(rows, counts):
    out, o = [], 0
    for c in counts.tolist():
        out.append(rows[o:o + c]); o += c
    return out


def main():
    rh.load_reference()
    import utils as U                                      # the reference's utils.py, unmodified
    rng = np.random.default_rng(0)
    raws, refs = [], []
    for name in ("synthetic_320x240_seed0.npz", "edge_cases.npz"):
        g = np.load(os.path.join(OUT, name))
        raws += split_rows(g["out_raw"], g["counts"]); refs += split_rows(g["out_refined"], g["counts"])
    targets, status, outs = [], [], []
    for i, (raw, ref) in enumerate(zip(raws, refs)):
        if raw.shape[0] == 0:
            t = np.array([[10.5, 20.25, 3.0]])              # labels but no corners: the caller skips the frame (inference.py:154)
        else:
            ids = np.unique(ref[:, 2])
            first = np.array([ref[ref[:, 2] == k][0] for k in ids])
            t = first.copy()
            t[:, :2] += rng.uniform(-1.5, 1.5, (len(ids), 2))
            t[:, :2] = t[:, :2].astype(np.float32)          # label_kpts.astype(np.float32) / up_scale, inference.py:148
            u = i % 5
            if u == 1 and len(ids) > 2:
                t = np.delete(t, 1, axis=0)                 # a predicted id without a label -> (None, None)
            elif u == 2:
                free = sorted(set(range(16)) - set(ids.astype(int).tolist()))
                if free:
                    t = np.concatenate([t, [[33.0, 44.5, float(free[0])]]], 0)     # a label nobody predicted: distance 0
                    t = t[np.argsort(t[:, 2], kind="stable")]
            elif u == 3 and i == 8:
                t = np.zeros((0, 3))                        # no labels
        targets.append(t)
        if raw.shape[0] == 0 or t.shape[0] == 0:
            status.append(0); outs.append(np.zeros(6)); continue
        with contextlib.redirect_stdout(io.StringIO()):
            m_raw, m_ref = U.pixel_error(raw, ref, t)
        if m_raw is None:
            status.append(0); outs.append(np.zeros(6)); continue
        d = U.compute_l2_distance(raw[:, :2], raw[:, 2], t[:, :2], t[:, 2])
        d_ref = U.compute_l2_distance(ref[:, :2], ref[:, 2], t[:, :2], t[:, 2])
        d_rr = U.compute_l2_distance(ref[:, :2], ref[:, 2], raw[:, :2], raw[:, 2])
        assert m_raw == d.mean() and m_ref == d_ref.mean()
        status.append(1); outs.append(np.array([d.mean(), d_ref.mean(), d_rr.mean(), d.max(), d_ref.max(), d_rr.max()]))
    tc = np.array([t.shape[0] for t in targets], np.int64)
    np.savez_compressed(os.path.join(OUT, "pixel_error_seed0.npz"), target_counts=tc,
                        targets=np.concatenate([t.reshape(-1, 3) for t in targets], 0).astype(np.float64),
                        status=np.array(status, np.int64), out=np.stack(outs).astype(np.float64),
                        meta=np.array("reference utils.pixel_error / compute_l2_distance on the golden raw/refined keypoints, numpy " + np.__version__))
    print("frames", len(targets), "status", status)
    print(np.round(np.stack(outs)[:6], 4))


if __name__ == "__main__":
    main()
