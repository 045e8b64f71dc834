"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU, fp32) in this container.

    python tools/make_golden.py

Every array stored here is an output of /root/reference code (imported through
tools/ref_harness.py), never of the oracle or of the CUDA engine.  The fixtures
pin `oracle/` (tests/test_oracle_golden.py) and, through it and directly, the
CUDA path on the GPU box where /root/reference does not exist.
"""
import os
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_harness as rh  # noqa: E402
from deepcharuco_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def ref_stages(ref, deepc, refinenet, gray_u8):
    """Stage outputs of the reference for one grayscale frame (inference.py:41-60, refinenet.py:104-114)."""
    from models.model_utils import pred_to_keypoints, extract_patches, pre_bgr_image
    img = torch.tensor(pre_bgr_image(gray_u8))
    loc, ids = deepc.infer_image(img)
    kpts, ids_found = pred_to_keypoints(loc, ids, 16)
    st = dict(loc=loc.numpy()[0], ids=ids.numpy()[0], kpts=kpts.numpy(), ids_found=ids_found.numpy())
    if ids_found.shape[0]:
        patches = extract_patches(img, kpts)
        with torch.no_grad():
            heat = refinenet.model(patches[:, None])[:, 0]
        refined, corners = refinenet.infer_patches(patches, kpts)
        st.update(patches=patches.numpy(), heat=heat.numpy(), refined=refined.numpy(), corners=corners.numpy())
    return st


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = rh.load_reference()
    deepc, refinenet = rh.load_reference_models("cpu")
    meta = dict(torch=torch.__version__, cv2=cv2.__version__, reference="JunkyByte/deepcharuco@37d569fc")

    # --- 1. the shipped sample image (benchmark.py:34, inference.py:181) -------------------------------
    bgr = cv2.imread(rh.SAMPLE_IMAGE)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    refined, _ = ref.infer_image(bgr, 16, deepc, refinenet)
    raw, _ = ref.infer_image(bgr, 16, deepc, None)
    st = ref_stages(ref, deepc, refinenet, gray)
    cam = np.array([[300, 0, 160], [0, 300, 120], [0, 0, 1]], np.float64)
    ret, rvec, tvec = ref.solve_pnp(refined, 5, 5, 0.01, cam, np.zeros(5))
    np.savez_compressed(os.path.join(OUT, "sample_image.npz"), bgr=bgr, gray=gray, out_refined=refined, out_raw=raw,
                        pnp_ret=np.array(ret), pnp_rvec=rvec, pnp_tvec=tvec, pnp_camera=cam,
                        **{"st_" + k: v for k, v in st.items()}, meta=np.array(str(meta)))
    print("sample:", refined.shape, raw.dtype, refined.dtype)

    # --- 2. seeded synthetic frames, seed 0 (the parity set) ----------------------------
    frames = synth.make_frames(16, 240, 320, seed=0)
    outs, raws, counts = [], [], []
    stages = []
    for f in frames:
        b = cv2.cvtColor(f, cv2.COLOR_GRAY2BGR)
        o, _ = ref.infer_image(b, 16, deepc, refinenet)
        r, _ = ref.infer_image(b, 16, deepc, None)
        outs.append(o.reshape(-1, 3)); raws.append(r.reshape(-1, 3)); counts.append(len(o))
        stages.append(ref_stages(ref, deepc, refinenet, f))
    n_logit = 3   # full logits for the first frames only (394 kB each)
    np.savez_compressed(
        os.path.join(OUT, "synthetic_320x240_seed0.npz"), frames=frames, counts=np.array(counts),
        out_refined=np.concatenate(outs, 0), out_raw=np.concatenate(raws, 0).astype(np.int64),
        loc=np.stack([s["loc"] for s in stages[:n_logit]]), ids=np.stack([s["ids"] for s in stages[:n_logit]]),
        patches=np.concatenate([s["patches"] for s in stages[:n_logit]], 0),
        heat=np.concatenate([s["heat"] for s in stages[:n_logit]], 0),
        corners=np.concatenate([s["corners"] for s in stages], 0),
        kpts=np.concatenate([s["kpts"] for s in stages], 0), ids_found=np.concatenate([s["ids_found"] for s in stages], 0),
        meta=np.array(str(meta)))
    print("synthetic 320x240: K =", counts)

    # --- 3. edge cases (SURVEY.md 8c invariants): K=0 frames, a crowded frame, a border corner ----------
    rng = np.random.default_rng(7)
    edge = {
        "zeros": np.zeros((240, 320), np.uint8),
        "white": np.full((240, 320), 255, np.uint8),
        "noise": rng.integers(0, 256, (240, 320)).astype(np.uint8),
    }
    # crowded: 12 small boards tiled 3x4 (duplicate ids, K >> 16)
    small = cv2.resize(synth.board_render(240), (80, 80), interpolation=cv2.INTER_AREA)
    edge["crowded"] = np.tile(small, (3, 4))
    # border: a board shifted so corners fall within 12 px of the frame edge (patch zero padding)
    big = np.full((240, 320), 90, np.uint8)
    b200 = cv2.resize(synth.board_render(240), (200, 200), interpolation=cv2.INTER_AREA)
    big[:160, :170] = b200[40:, 30:]
    edge["border"] = big
    names = sorted(edge)
    e_out, e_raw, e_cnt = [], [], []
    for nme in names:
        b = cv2.cvtColor(edge[nme], cv2.COLOR_GRAY2BGR)
        o, _ = ref.infer_image(b, 16, deepc, refinenet)
        r, _ = ref.infer_image(b, 16, deepc, None)
        e_out.append(np.asarray(o, np.float64).reshape(-1, 3)); e_raw.append(np.asarray(r, np.int64).reshape(-1, 3))
        e_cnt.append(len(o))
        print("edge", nme, "K =", len(o))
    np.savez_compressed(os.path.join(OUT, "edge_cases.npz"), names=np.array(names),
                        frames=np.stack([edge[n] for n in names]), counts=np.array(e_cnt),
                        out_refined=np.concatenate(e_out, 0), out_raw=np.concatenate(e_raw, 0), meta=np.array(str(meta)))

    # --- 4. one 640x480 frame (config 5 shape) ----------------------------------------------------------
    f640 = synth.make_frames(2, 480, 640, seed=3)
    o640, c640 = [], []
    for f in f640:
        o, _ = ref.infer_image(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR), 16, deepc, refinenet)
        o640.append(o.reshape(-1, 3)); c640.append(len(o))
    np.savez_compressed(os.path.join(OUT, "synthetic_640x480_seed3.npz"), frames=f640, counts=np.array(c640),
                        out_refined=np.concatenate(o640, 0), meta=np.array(str(meta)))
    print("synthetic 640x480: K =", c640)


if __name__ == "__main__":
    main()
