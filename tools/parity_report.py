"""Parity report: CUDA engine (both conv implementations) vs the oracle on seeded synthetic frames.

    python tools/parity_report.py [--frames 128] [--seed 1] [--out gpurun_out/parity.json]

For every frame: kept-cell set, ids, raw pixels (bit-exact expected), refined corners (<= 1e-3 px expected); every
mismatch is classified by the oracle's own top-1/top-2 margin.  Also reports max |delta| of logits / heat maps.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepcharuco_b200 as dc  # noqa: E402
import oracle  # noqa: E402
import parity  # noqa: E402
from deepcharuco_b200 import _native as N, synth, weights_io as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--out", default="gpurun_out/parity.json")
a = ap.parse_args()

states = (W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET))
frames = synth.make_frames(a.frames, 240, 320, seed=a.seed)
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.time()
orc = [oracle.pipeline.infer_gray(states[0], states[1], f, return_stages=True) for f in frames]
t_or = time.time() - t0
report = dict(frames=a.frames, seed=a.seed, oracle_seconds=t_or, torch=torch.__version__, impls={})
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, 16, "cuda:0")
for name, impl in (("tcgen05_f16x2", N.CONV_TCGEN05), ("ffma_fp32", N.CONV_FFMA)):
    deepc._ctx.set_conv_impl(impl)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    reps = [parity.compare_frame(states, f, r, w) for f, r, w in zip(frames, refined, raw)]
    tot = parity.summarise(reps)
    # logits / heat deltas on the first 16 frames through the stage entry points
    eng = deepc._ctx.engine(240, 320, max_batch=a.frames)
    n = min(16, a.frames)
    fr = torch.from_numpy(frames[:n]).cuda()
    loc = torch.empty((n, 65, 30, 40), device="cuda"); ids = torch.empty((n, 17, 30, 40), device="cuda")
    N.check(N.lib().dcu_detector_forward(eng.handle, fr.data_ptr(), n, loc.data_ptr(), ids.data_ptr(), None))
    torch.cuda.synchronize()
    dl = max(float(np.abs(loc[i].cpu().numpy() - orc[i][1]["loc"][0]).max()) for i in range(n))
    di = max(float(np.abs(ids[i].cpu().numpy() - orc[i][1]["ids"][0]).max()) for i in range(n))
    patches = np.concatenate([orc[i][1]["patches"] for i in range(n) if "patches" in orc[i][1]], 0)
    kp = np.concatenate([orc[i][1]["kpts"] for i in range(n) if "patches" in orc[i][1]], 0).astype(np.int32)
    heat_o = np.concatenate([orc[i][1]["heat"] for i in range(n) if "patches" in orc[i][1]], 0)
    P = patches.shape[0]
    dp, dk = torch.from_numpy(patches).cuda(), torch.from_numpy(kp).cuda()
    corners = torch.empty((P, 2), dtype=torch.int32, device="cuda"); ref = torch.empty((P, 2), device="cuda")
    heat = torch.empty((P, 64, 64), device="cuda")
    N.check(N.lib().dcu_refine_forward(eng.handle, dp.data_ptr(), dk.data_ptr(), 2, P, corners.data_ptr(), ref.data_ptr(), heat.data_ptr(), None))
    torch.cuda.synchronize()
    dh = float(np.abs(heat.cpu().numpy() - heat_o).max())
    tot.update(max_abs_dloc=dl, max_abs_dids=di, max_abs_dheat=dh, heat_patches=int(P))
    report["impls"][name] = tot
    print(name, json.dumps(tot), flush=True)
# oracle margins, for context
m_heat = []
for res, st in orc:
    if "heat" in st:
        h = st["heat"].reshape(st["heat"].shape[0], -1)
        top2 = np.sort(h, axis=1)[:, -2:]
        m_heat += (top2[:, 1] - top2[:, 0]).tolist()
m = np.array(m_heat)
report["oracle_heat_margin"] = dict(n=int(m.size), min=float(m.min()), q01=float(np.quantile(m, 0.01)), q05=float(np.quantile(m, 0.05)),
                                    median=float(np.median(m)))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(report, open(a.out, "w"), indent=1)
print("oracle heat margins:", report["oracle_heat_margin"])
