"""Throughput of every BASELINE.json config on ONE B200 (config 5 = one GPU's 256-frame shard of the 2048 x 640x480 batch).
Device-resident inputs, CUDA events on the launching stream, 3 warm-ups + `--iters` timed passes.  Not the driver's bench
(bench.py is); this is the per-config evidence table for profiles/.

    python tools/bench_configs.py [--iters 5] [--out gpurun_out/configs.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepcharuco_b200 import _native as N, synth, weights_io as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/configs.json")
a = ap.parse_args()
sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)
L = N.lib()
res = {}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# config 2: batch=64 320x240, detector forward only
eng = N.Engine(sd, sr, 240, 320, 16, 0, max_batch=256, max_patches=16384)
frames = torch.from_numpy(synth.tile_frames(synth.make_frames(64, seed=1), 256)).cuda()
loc = torch.empty((64, 65, 30, 40), device="cuda"); ids = torch.empty((64, 17, 30, 40), device="cuda")
ms = timed(lambda: N.check(L.dcu_detector_forward(eng.handle, frames.data_ptr(), 64, loc.data_ptr(), ids.data_ptr(), None)), a.iters)
res["config2_detector_b64_320x240"] = dict(ms=ms, frames_per_s=64 / ms * 1e3, tflops_alg=64 * eng.detector_flops_per_frame() / ms / 1e9)
# config 3: batch=256 full pipeline
ms = timed(lambda: eng.infer_batch_device(frames.data_ptr(), 256, 16, True, None), a.iters)
k = int(eng._dev_out["total"].item())
res["config3_full_b256_320x240"] = dict(ms=ms, frames_per_s=256 / ms * 1e3, corners=k,
                                         tflops_alg=(256 * eng.detector_flops_per_frame() + k * eng.refine_flops_per_patch()) / ms / 1e9)
# board pose for the 256 frames of config 3, GPU batch (on the device-resident results) vs the reference's
# per-frame cv2.solvePnP loop on one host core (inference.py:15-29)
import time
import deepcharuco_b200 as dc
Kc = np.array([[300.0, 0, 160], [0, 300.0, 120], [0, 0, 1]])
ms = timed(lambda: eng.solve_pnp_batch_device(256, 5, 5, 0.01, Kc, np.zeros(5), True, None), 20)
host_rows = dc.inference._rows_to_frames(*eng.infer_batch_host(frames.cpu().numpy(), 16, True))
t0 = time.time()
for kp in host_rows:
    dc.solve_pnp(kp, 5, 5, 0.01, Kc, np.zeros(5)) if kp.size else None
cv_ms = (time.time() - t0) * 1e3
res["pnp_b256"] = dict(gpu_ms=ms, gpu_frames_per_s=256 / ms * 1e3, cv2_host_loop_ms=cv_ms, cv2_frames_per_s=256 / cv_ms * 1e3)
# config 4: RefineNet microbench, 16384 real patches (the engine's own gather output, cycled)
o = eng._dev_out
patches_src = torch.empty((k, 24, 24), device="cuda")
counts = torch.empty(256, dtype=torch.int32, device="cuda"); offs = torch.empty(256, dtype=torch.int32, device="cuda")
tot = torch.zeros(1, dtype=torch.int32, device="cuda"); kp = torch.empty((16384, 4), dtype=torch.int32, device="cuda")
loc256 = torch.empty((256, 65, 30, 40), device="cuda"); ids256 = torch.empty((256, 17, 30, 40), device="cuda")
N.check(L.dcu_detector_forward(eng.handle, frames.data_ptr(), 256, loc256.data_ptr(), ids256.data_ptr(), None))
pt = torch.empty((16384, 24, 24), device="cuda")
N.check(L.dcu_decode_gather(eng.handle, loc256.data_ptr(), ids256.data_ptr(), frames.data_ptr(), 256, 16, 0, counts.data_ptr(),
                            offs.data_ptr(), tot.data_ptr(), kp.data_ptr(), pt.data_ptr(), None))
torch.cuda.synchronize()
k2 = int(tot.item())
reps = (16384 + k2 - 1) // k2
pt16 = pt[:k2].repeat(reps, 1, 1)[:16384].contiguous()
kp16 = kp[:k2].repeat(reps, 1)[:16384].contiguous()
corners = torch.empty((16384, 2), dtype=torch.int32, device="cuda"); refined = torch.empty((16384, 2), device="cuda")
ms = timed(lambda: N.check(L.dcu_refine_forward(eng.handle, pt16.data_ptr(), kp16.data_ptr(), 4, 16384, corners.data_ptr(),
                                                refined.data_ptr(), None, None)), a.iters)
res["config4_refinenet_16384_patches"] = dict(ms=ms, patches_per_s=16384 / ms * 1e3, tflops_alg=16384 * eng.refine_flops_per_patch() / ms / 1e9)
# decode + gather alone (HBM-bound kernel): 256 frames
ms = timed(lambda: N.check(L.dcu_decode_gather(eng.handle, loc256.data_ptr(), ids256.data_ptr(), frames.data_ptr(), 256, 16, 0,
                                               counts.data_ptr(), offs.data_ptr(), tot.data_ptr(), kp.data_ptr(), pt.data_ptr(), None)), 20)
nbytes = 256 * 82 * 1200 * 4 + k2 * (2304 + 2304 + 16)
res["decode_gather_b256_320x240"] = dict(ms=ms, frames_per_s=256 / ms * 1e3, gb_per_s_alg=nbytes / ms / 1e6, corners=k2)
eng.close()
# config 5 (per-GPU shard): 256 frames 640x480, 4 boards per frame
eng5 = N.Engine(sd, sr, 480, 640, 16, 0, max_batch=256, max_patches=32768)
f5 = torch.from_numpy(synth.tile_frames(synth.make_frames(16, 480, 640, seed=1), 256)).cuda()
ms = timed(lambda: eng5.infer_batch_device(f5.data_ptr(), 256, 16, True, None), max(2, a.iters // 2))
k5 = int(eng5._dev_out["total"].item())
res["config5_full_b256_640x480_per_gpu"] = dict(ms=ms, frames_per_s=256 / ms * 1e3, corners=k5,
                                                 tflops_alg=(256 * eng5.detector_flops_per_frame() + k5 * eng5.refine_flops_per_patch()) / ms / 1e9)
eng5.close()
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps(res, indent=1))
