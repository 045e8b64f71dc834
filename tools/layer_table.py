"""Per-launch timing table of one bench step (CUDA events around every launch, dcu_profile_records).

    python tools/layer_table.py --batch 256 [--json out.json]

Prints, per distinct layer shape: launches, total ms, algorithmic TFLOP/s, tile efficiency (useful / issued pixels of the
kernel's CTA tiling) and issued tensor TFLOP/s (3 products per MAC / tile efficiency).  Profiling serialises the conv1a side
stream, so the sum is a little above an unprofiled step."""
import argparse
import json
import math
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepcharuco_b200 as dc  # noqa: E402
from deepcharuco_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--json", default="")
a = ap.parse_args()
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, 16, "cuda:0")
eng = deepc._ctx.engine(240, 320, max_batch=a.batch, max_patches=64 * a.batch)
frames = torch.from_numpy(synth.tile_frames(synth.make_frames(64, seed=1), a.batch)).cuda()
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    eng.infer_batch_device(frames.data_ptr(), a.batch, 16, True, s)
torch.cuda.synchronize()
eng.profile_enable(True)
REP = 3
for _ in range(REP):
    eng.infer_batch_device(frames.data_ptr(), a.batch, 16, True, s)
rec = eng.profile_records()
eng.profile_enable(False)
names = {0: "conv3x3_tc", 1: "conv_first", 2: "heads_1x1", 3: "decode_gather", 4: "refine_finalize"}
rows = OrderedDict()
for r in rec:
    cls, ms, work, cin, cout, ho, wo, n = r
    key = (int(cls), int(cin), int(cout), int(ho), int(wo))
    d = rows.setdefault(key, {"launches": 0, "ms": 0.0, "work": 0.0, "n": 0})
    d["launches"] += 1; d["ms"] += ms; d["work"] += work; d["n"] += int(n)
out = []
tot = 0.0
for (cls, cin, cout, ho, wo), d in rows.items():
    ms = d["ms"] / REP
    tot += ms
    row = {"kernel": names[cls], "cin": cin, "cout": cout, "hout": ho, "wout": wo, "launches": d["launches"] // REP,
           "ms_per_step": round(ms, 4)}
    if cls in (0, 1, 2) and d["work"] > 0:
        row["alg_tflops"] = round(d["work"] / REP / ms / 1e9, 1)
    if cls == 3:
        row["alg_gbs"] = round(d["work"] / REP / ms / 1e6, 1)
    out.append(row)
for row in out:
    row["share"] = round(row["ms_per_step"] / tot, 4)
    print(row)
print("sum of launches: %.3f ms per step" % tot)
if a.json:
    json.dump({"batch": a.batch, "sum_ms": tot, "rows": out}, open(a.json, "w"), indent=1)
