"""One warm step + one profiled step of the bench workload, for `ncu --profile-from-start off` (never a bench value).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --batch 256
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepcharuco_b200 as dc  # noqa: E402
from deepcharuco_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--conv", default="")
a = ap.parse_args()
if a.conv:
    os.environ["DCU_CONV_IMPL"] = a.conv
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, 16, "cuda:0")
eng = deepc._ctx.engine(240, 320, max_batch=a.batch, max_patches=64 * a.batch)
frames = torch.from_numpy(synth.tile_frames(synth.make_frames(64, seed=1), a.batch)).cuda()
s = torch.cuda.current_stream().cuda_stream
eng.infer_batch_device(frames.data_ptr(), a.batch, 16, True, s)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.infer_batch_device(frames.data_ptr(), a.batch, 16, True, s)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step: batch", a.batch, "launches so far", eng.launch_count())
