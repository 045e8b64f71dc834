"""Role-level cycle accounting of the CTA-pair tcgen05 conv kernel per layer shape (profiling aid; not a bench).

Needs a library whose conv_tc2.cu was compiled with -DDCU_TC2_STATS (the shipped build compiles the counters out)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepcharuco_b200 import _native as N, weights_io as W  # noqa: E402

sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)
L = N.lib()
# (net, layer, cin, h, w, cout, out_h, out_w, images, name); h, w = the layer's input size as the reference sees it
CASES = [(0, 1, 64, 240, 320, 64, 120, 160, 64, "det conv1b 64->64 @240x320 x64   <64,3,0,1> resident weights"),
         (0, 2, 64, 120, 160, 64, 120, 160, 64, "det conv2a 64->64 @120x160 x64"),
         (0, 5, 128, 60, 80, 128, 30, 40, 64, "det conv3b 128->128 @60x80 x64   <128,3,0,0,0,1> one m-tile"),
         (0, 8, 128, 30, 40, 512, 30, 40, 64, "det convPa|Da 128->512 @30x40 x64"),
         (1, 6, 128, 16, 16, 128, 16, 16, 2048, "ref conv4a 128->128 (8x8 up) x2048   <64,3,1,0> upsample-fused"),
         (1, 8, 128, 32, 32, 64, 32, 32, 2048, "ref conv5a 128->64 (16x16 up) x2048"),
         (1, 10, 64, 64, 64, 64, 64, 64, 1024, "ref convPa 64->64 (32x32 up) x1024   <64,3,1,1>")]
names = ["mma_total", "mma_wait_halo", "mma_wait_weights", "mma_wait_epilogue", "-", "-", "epi_total", "epi_wait_mma"]
eng = N.Engine(sd, sr, 240, 320, 16, 0, max_batch=64, max_patches=4096)
for net, layer, cin, h, w, cout, oh, ow, n, name in CASES:
    x = torch.rand((n, cin, h, w), device="cuda")
    out = torch.empty((n, cout, oh, ow), device="cuda")
    for rep in range(2):
        L.dcu_debug_tc_stats(eng.handle, 1, None)
        N.check(L.dcu_debug_conv_layer(eng.handle, net, layer, N.CONV_TCGEN05, x.data_ptr(), n, h, w, out.data_ptr(), None))
    buf = (C.c_uint64 * 8)()
    L.dcu_debug_tc_stats(eng.handle, 0, buf)
    v = np.array(list(buf), dtype=np.float64)
    if v[0] == 0:
        print(name, ": no counters (library built without -DDCU_TC2_STATS, see conv_tc2.cu)")
        continue
    ctas = 74          # counters come from the leader CTA of every pair
    print(name)
    print("   per-leader avg kcycles:", {k: round(x / ctas / 1e3, 1) for k, x in zip(names, v) if k != "-"})
    print("   MMA warp: wait_halo %.3f wait_weights %.3f wait_epilogue %.3f issue+other %.3f | epilogue warp waits for MMAs %.3f of its time" % (
        v[1] / v[0], v[2] / v[0], v[3] / v[0], 1 - (v[1] + v[2] + v[3]) / v[0], v[7] / max(v[6], 1)))
