"""Role-level cycle accounting of the tcgen05 conv kernel on one layer shape (profiling aid; not a bench)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepcharuco_b200 import _native as N, weights_io as W  # noqa: E402

sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)
eng = N.Engine(sd, sr, 240, 320, 16, 0, max_batch=32, max_patches=2048)
L = N.lib()
CASES = [(0, 1, 64, 240, 320, 64, 120, 160, 32, "det conv1b 64->64 @240x320 x32"),
         (0, 2, 64, 120, 160, 64, 120, 160, 32, "det conv2a 64->64 @120x160 x32"),
         (0, 5, 128, 60, 80, 128, 30, 40, 32, "det conv3b 128->128 @60x80 x32"),
         (0, 8, 128, 30, 40, 512, 30, 40, 32, "det convPa|Da 128->512 @30x40 x32"),
         (1, 10, 64, 64, 64, 64, 64, 64, 256, "ref convPa 64->64 @64x64 x256"),
         (1, 4, 128, 8, 8, 128, 8, 8, 1024, "ref conv3a 128->128 @8x8 x1024")]
names = ["mma_total", "mma_wait_halo", "mma_wait_weights", "mma_wait_epilogue", "split_total", "split_wait_tma", "epi_total", "epi_wait_mma"]
for net, layer, cin, h, w, cout, oh, ow, n, name in CASES:
    x = torch.rand((n, cin, h, w), device="cuda")
    out = torch.empty((n, cout, oh, ow), device="cuda")
    for rep in range(2):
        L.dcu_debug_tc_stats(eng.handle, 1, None)
        N.check(L.dcu_debug_conv_layer(eng.handle, net, layer, N.CONV_TCGEN05, x.data_ptr(), n, h, w, out.data_ptr(), None))
    buf = (C.c_uint64 * 8)()
    L.dcu_debug_tc_stats(eng.handle, 0, buf)
    v = np.array(list(buf), dtype=np.float64)
    ctas = 148
    print(name)
    print("   per-CTA avg kcycles:", {k: round(x / ctas / 1e3, 1) for k, x in zip(names, v)})
    print("   fractions of MMA-warp time: wait_halo %.2f wait_weights %.2f wait_epilogue %.2f issue %.2f" % (
        v[1] / v[0], v[2] / v[0], v[3] / v[0], 1 - (v[1] + v[2] + v[3]) / v[0]))
