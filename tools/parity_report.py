"""Parity report: CUDA engine vs the oracle on seeded synthetic frames.  Run on the GPU box.

    python tools/parity_report.py [--frames 256] [--seed 1] [--impls tcgen05,ffma] [--size 320x240] [--out gpurun_out/parity.json]

Frames are generated and checked in chunks of 256 (chunk c uses seed + c), so --frames 4096 is the >= 50 000-corner run.
For every frame: kept-cell set, ids, raw pixels (bit-exact expected), refined corners (<= 1e-3 px expected); every mismatch is
listed with the oracle's own top-1 margin at the engine's choice (tests/parity.py).  Also reports max |delta| of logits / heat
maps (first chunk) and the distribution of the oracle's margins, so that the flip count can be read against it.
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
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--impls", default="tcgen05,ffma")
ap.add_argument("--size", default="320x240")
ap.add_argument("--out", default="gpurun_out/parity.json")
a = ap.parse_args()
Wd, Hd = [int(v) for v in a.size.split("x")]
IMPLS = {"tcgen05": ("tcgen05_f16x2", N.CONV_TCGEN05), "ffma": ("ffma_fp32", N.CONV_FFMA)}
impls = [IMPLS[k] for k in a.impls.split(",")]

states = (W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET))
torch.set_num_threads(os.cpu_count() or 1)
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, 16, "cuda:0")
report = dict(frames=a.frames, seed=a.seed, size=a.size, torch=torch.__version__, tie_tolerances=dict(heat=parity.HEAT_TIE_TOL, loc=parity.LOC_TIE_TOL),
              impls={name: None for name, _ in impls})
reps = {name: [] for name, _ in impls}
m_heat, m_loc = [], []
t_or = 0.0
CH = 256
for c0 in range(0, a.frames, CH):
    n = min(CH, a.frames - c0)
    frames = synth.make_frames(n, Hd, Wd, seed=a.seed + c0 // CH)
    cache = {}
    t0 = time.time()
    for i, f in enumerate(frames):
        cache[i] = parity.oracle_stages(states, f)
    t_or += time.time() - t0
    for res, st in cache.values():
        if "heat" in st:
            h = st["heat"].reshape(st["heat"].shape[0], -1)
            top2 = np.sort(h, axis=1)[:, -2:]
            m_heat += (top2[:, 1] - top2[:, 0]).tolist()
            for (x, y) in st["kpts"]:
                col = np.sort(st["loc"][0, :, int(y) // 8, int(x) // 8])
                m_loc.append(float(col[-1] - col[-2]))
    for name, impl in impls:
        deepc._ctx.set_conv_impl(impl)
        refined = dc.infer_batch(frames, 16, deepc, refinenet)
        raw = dc.infer_batch(frames, 16, deepc, None)
        reps[name] += [parity.compare_frame(states, f, r, w, cache=cache, key=i) for i, (f, r, w) in enumerate(zip(frames, refined, raw))]
        if c0 == 0:
            # logits / heat deltas on the first 16 frames through the stage entry points
            eng = deepc._ctx.engine(Hd, Wd, max_batch=n)
            k = min(16, n)
            fr = torch.from_numpy(frames[:k]).cuda()
            loc = torch.empty((k, 65, Hd // 8, Wd // 8), device="cuda"); ids = torch.empty((k, 17, Hd // 8, Wd // 8), device="cuda")
            N.check(N.lib().dcu_detector_forward(eng.handle, fr.data_ptr(), k, loc.data_ptr(), ids.data_ptr(), None))
            torch.cuda.synchronize()
            dl = max(float(np.abs(loc[i].cpu().numpy() - cache[i][1]["loc"][0]).max()) for i in range(k))
            di = max(float(np.abs(ids[i].cpu().numpy() - cache[i][1]["ids"][0]).max()) for i in range(k))
            with_p = [i for i in range(k) if "patches" in cache[i][1]]
            patches = np.concatenate([cache[i][1]["patches"] for i in with_p], 0)
            kp = np.concatenate([cache[i][1]["kpts"] for i in with_p], 0).astype(np.int32)
            heat_o = np.concatenate([cache[i][1]["heat"] for i in with_p], 0)
            P = patches.shape[0]
            dp, dk = torch.from_numpy(patches).cuda(), torch.from_numpy(kp).cuda()
            corners = torch.empty((P, 2), dtype=torch.int32, device="cuda"); ref = torch.empty((P, 2), device="cuda")
            heat = torch.empty((P, 64, 64), device="cuda")
            N.check(N.lib().dcu_refine_forward(eng.handle, dp.data_ptr(), dk.data_ptr(), 2, P, corners.data_ptr(), ref.data_ptr(), heat.data_ptr(), None))
            torch.cuda.synchronize()
            dh = float(np.abs(heat.cpu().numpy() - heat_o).max())
            report["impls"][name] = dict(max_abs_dloc=dl, max_abs_dids=di, max_abs_dheat=dh, heat_patches=int(P))
    print(f"chunk {c0 // CH}: " + "  ".join(f"{name}: K={parity.summarise(reps[name])['K']} flips={len(parity.summarise(reps[name])['flips'])}" for name, _ in impls), flush=True)
for name, _ in impls:
    tot = parity.summarise(reps[name])
    tot["unexplained"] = (tot["raw_px"] - tot["raw_px_explained"]) + (tot["heat_flip"] - tot["heat_flip_explained"])
    report["impls"][name] = dict(report["impls"][name] or {}, **tot)
    print(name, json.dumps({k: v for k, v in report["impls"][name].items() if k != "flips"}), flush=True)
    for fl in tot["flips"]:
        print("   flip:", fl)
report["oracle_seconds"] = t_or
m = np.array(m_heat); ml = np.array(m_loc)
report["oracle_heat_margin"] = dict(n=int(m.size), min=float(m.min()), q001=float(np.quantile(m, 0.001)), q01=float(np.quantile(m, 0.01)),
                                    q05=float(np.quantile(m, 0.05)), median=float(np.median(m)), below_1e_5=int((m < 1e-5).sum()),
                                    below_5e_5=int((m < 5e-5).sum()))
report["oracle_loc_margin"] = dict(n=int(ml.size), min=float(ml.min()), q001=float(np.quantile(ml, 0.001)), q01=float(np.quantile(ml, 0.01)),
                                   median=float(np.median(ml)), below_5e_3=int((ml < 5e-3).sum()), below_2e_2=int((ml < 2e-2).sum()))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(report, open(a.out, "w"), indent=1)
print("oracle heat margins:", report["oracle_heat_margin"])
print("oracle loc margins:", report["oracle_loc_margin"])
