"""CPU emulation of the operand split the tensor-core convolution uses, on one trained layer: the three products
a_hi*w_hi + a_hi*w_lo + a_lo*w_hi of the fp16 hi/lo split (power-of-two weight scaling, fp32 accumulate) reproduce the fp64
convolution as well as a plain fp32 convolution does, while a single fp16 pass is ~1000x worse -- which is why the engine pays for
three products (single-pass fp16 / TF32 flips 0.4 % of the refined corners).  The whole-network emulation is
tools/emulate_split.py."""
import numpy as np
import torch
import torch.nn.functional as F


def _split_f16(x, scale=1.0):
    xs = x * scale
    hi = xs.half().float()
    lo = (xs - hi).half().float()
    return hi, lo


def test_three_product_fp16_split_is_fp32_equivalent(states):
    sd = states[0]
    w = torch.from_numpy(sd["conv3b.weight"])                       # 128 -> 128, 3x3
    g = torch.Generator().manual_seed(0)
    x = torch.relu(torch.randn(1, 128, 30, 40, generator=g)) * 3.0  # post-ReLU activations of realistic magnitude
    truth = F.conv2d(x.double(), w.double(), padding=1)
    scale = float(2.0 ** np.floor(np.log2(32768.0 / float(w.abs().max()))))
    xh, xl = _split_f16(x)
    wh, wl = _split_f16(w, scale)
    main = F.conv2d(xh.double(), wh.double(), padding=1)
    small = F.conv2d(xh.double(), wl.double(), padding=1) + F.conv2d(xl.double(), wh.double(), padding=1)
    split3 = (main.float() + small.float()) / scale                 # the kernel's epilogue: fp32 sum of the two accumulators, * 2^-s
    fp32 = F.conv2d(x, w, padding=1)
    single = F.conv2d(xh.double(), wh.double(), padding=1).float() / scale
    ref = float(truth.abs().max())
    e_split = float((split3.double() - truth).abs().max()) / ref
    e_fp32 = float((fp32.double() - truth).abs().max()) / ref
    e_single = float((single.double() - truth).abs().max()) / ref
    assert e_split < 5e-7 and e_split < 4 * e_fp32 + 1e-7, (e_split, e_fp32)
    assert e_single > 100 * e_split, (e_single, e_split)
    # the power-of-two scaling keeps the low halves of the weights in fp16's normal range
    assert float(wl[wl != 0].abs().min()) >= 2.0 ** -24 and float((w * scale).abs().max()) < 65504
