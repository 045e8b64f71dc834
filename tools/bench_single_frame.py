"""The reference's own benchmark loop (src/benchmark.py:38-53) on the drop-in surface: 5 warm-ups, then 500 calls of
infer_image(sample, draw_pred=False), wall clock, n / elapsed.  BASELINE config 1 shape (one 320x240 frame per call).

    python tools/bench_single_frame.py [--cpu]      # --cpu: also time the oracle port on the host cores
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepcharuco_b200 as dc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--n", type=int, default=500)
a = ap.parse_args()
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "sample_image.npz"))
img = g["bgr"]
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda")
for _ in range(5):
    kp, _ = dc.infer_image(img, 16, deepc, refinenet, draw_pred=False)
t = time.time()
for _ in range(a.n):
    kp, _ = dc.infer_image(img, 16, deepc, refinenet, draw_pred=False)
dt = time.time() - t
out = dict(workload="IMG_7412.png 320x240, one frame per call, full pipeline (src/benchmark.py loop)", calls=a.n,
           fps=a.n / dt, ms_per_call=dt / a.n * 1e3, corners=int(kp.shape[0]), matches_golden=bool(np.abs(kp - g["out_refined"]).max() <= 1e-3))
if a.cpu:
    import torch
    import oracle
    from deepcharuco_b200 import weights_io as W
    sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(3):
        oracle.infer_image(sd, sr, img)
    t = time.time()
    m = 60
    for _ in range(m):
        oracle.infer_image(sd, sr, img)
    out["cpu_oracle_fps"] = m / (time.time() - t)
    out["cpu_cores"] = os.cpu_count()
print(json.dumps(out))
