"""CPU model of the tcgen05 accumulation error of the split-precision convolution (build-container tool, slow).

The kernel (csrc/conv_tc2.cu) issues, per (16-channel chunk, tap), one MMA  a_hi x [w_hi | w_lo]  and one  a_lo x w_hi  with
fp32 accumulators in tensor memory.  fp16 x fp16 products are exact in fp32 and the 16 products of one MMA are summed in a
wide adder; what is NOT exact is the addition to the running fp32 accumulator.  This tool models that addition as
   acc <- round_mode(acc + exact_sum_of_16_products)           round_mode in {rz (truncate), rn}
and optionally drains the accumulators into an fp32 register every `seg` MMAs (two-level accumulation, RN adds), to
predict how the logit / heat-map error depends on the hardware rounding mode and on the chain length.

    python tools/emulate_mma_trunc.py [--frames 2] [--modes rz,rn] [--segs 0,9,3]

Prints max |dloc|, |dids|, |dheat| against the fp32 oracle for every (mode, seg).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle  # noqa: E402
from deepcharuco_b200 import synth, weights_io as W  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)


def split_f16(x):
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def to_f32(x64, mode):
    """double -> float32 with round-to-nearest ('rn') or truncation toward zero ('rz')."""
    y = x64.float()
    if mode == "rn":
        return y
    over = y.double().abs() > x64.abs()
    y = torch.where(over, torch.nextafter(y, torch.zeros_like(y)), y)
    return y


LO8_SHIFT = None          # set to p: the a_lo x w_hi product runs on e4m3 operands (a_lo * 2^p, w_hi * 2^-p), --lo8 p


def q8(x):
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def conv_model(x, w, pad, mode, seg, taps_collapsed=None):
    """x (N,C,H,W) fp32, w (O,C,3,3) fp32 -> conv output (N,O,H',W') fp32 before bias, per the kernel's MMA order."""
    O, C = w.shape[0], w.shape[1]
    if isinstance(seg, dict):            # per-layer policy: MMAs per segment by input channel count
        seg = seg.get(C, 0)
    mx = float(w.abs().max())
    s = 2.0 ** np.floor(np.log2(32768.0 / mx))
    xh, xl = split_f16(x)
    wh, wl = split_f16(w * s)
    wx = wh
    if LO8_SHIFT is not None:
        xl = q8(xl * 2.0 ** LO8_SHIFT) * 2.0 ** -LO8_SHIFT
        wx = q8((w * s) * 2.0 ** -LO8_SHIFT) * 2.0 ** LO8_SHIFT
    wx = wx.double()
    xh = F.pad(xh, (pad,) * 4).double()
    xl = F.pad(xl, (pad,) * 4).double()
    wh, wl = wh.double(), wl.double()
    N, _, Hp, Wp = xh.shape
    Ho, Wo = Hp - 2, Wp - 2
    main = torch.zeros((N, O, Ho, Wo))
    small = torch.zeros((N, O, Ho, Wo))
    reg = torch.zeros((N, O, Ho, Wo))
    count = 0
    for q in range(C // 16):
        cs = slice(q * 16, q * 16 + 16)
        for ky in range(3):
            for kx in range(3):
                ah = xh[:, cs, ky:ky + Ho, kx:kx + Wo]
                al = xl[:, cs, ky:ky + Ho, kx:kx + Wo]
                pm = torch.einsum("nchw,oc->nohw", ah, wh[:, cs, ky, kx])
                ps1 = torch.einsum("nchw,oc->nohw", ah, wl[:, cs, ky, kx])
                ps2 = torch.einsum("nchw,oc->nohw", al, wx[:, cs, ky, kx])
                main = to_f32(main.double() + pm, mode)
                small = to_f32(small.double() + ps1, mode)
                small = to_f32(small.double() + ps2, mode)
                count += 1
                if seg and count % seg == 0:
                    reg = reg + (main + small)
                    main.zero_(); small.zero_()
    return (reg + (main + small)) * np.float32(1.0 / s)


def cbr(x, st, name, pad, mode, seg):
    w = torch.from_numpy(st[name + ".weight"]); b = torch.from_numpy(st[name + ".bias"])
    if w.shape[1] == 1 or mode == "fp32":
        y = F.conv2d(x, w, b, padding=pad)
    else:
        y = conv_model(x, w, pad, mode, seg) + b.view(1, -1, 1, 1)
    bn = "bn" + name[4:]
    y = F.batch_norm(y, torch.from_numpy(st[bn + ".running_mean"]), torch.from_numpy(st[bn + ".running_var"]),
                     torch.from_numpy(st[bn + ".weight"]), torch.from_numpy(st[bn + ".bias"]), False, 0.1, 1e-5)
    return F.relu(y)


def head_1x1(x, w, b, mode):
    """1x1 heads on the single-CTA kernel: 16 chunks of 16 channels, one MMA pair each."""
    if mode == "fp32":
        return F.conv2d(x, w, b)
    O, C = w.shape[0], w.shape[1]
    mx = float(w.abs().max())
    s = 2.0 ** np.floor(np.log2(32768.0 / mx))
    xh, xl = split_f16(x); wh, wl = split_f16(w[:, :, 0, 0] * s)
    xh, xl, wh, wl = xh.double(), xl.double(), wh.double(), wl.double()
    main = torch.zeros((x.shape[0], O) + tuple(x.shape[2:])); small = torch.zeros_like(main)
    for q in range(C // 16):
        cs = slice(q * 16, q * 16 + 16)
        main = to_f32(main.double() + torch.einsum("nchw,oc->nohw", xh[:, cs], wh[:, cs]), mode)
        small = to_f32(small.double() + torch.einsum("nchw,oc->nohw", xh[:, cs], wl[:, cs]), mode)
        small = to_f32(small.double() + torch.einsum("nchw,oc->nohw", xl[:, cs], wh[:, cs]), mode)
    return (main + small) * np.float32(1.0 / s) + b.view(1, -1, 1, 1)


def det(x, mode, seg):
    st = sd
    x = cbr(x, st, "conv1a", 1, mode, seg); x = cbr(x, st, "conv1b", 1, mode, seg); x = F.max_pool2d(x, 2, 2)
    x = cbr(x, st, "conv2a", 1, mode, seg); x = cbr(x, st, "conv2b", 1, mode, seg); x = F.max_pool2d(x, 2, 2)
    x = cbr(x, st, "conv3a", 1, mode, seg); x = cbr(x, st, "conv3b", 1, mode, seg); x = F.max_pool2d(x, 2, 2)
    x = cbr(x, st, "conv4a", 1, mode, seg); x = cbr(x, st, "conv4b", 1, mode, seg)
    pa = cbr(x, st, "convPa", 1, mode, seg); da = cbr(x, st, "convDa", 1, mode, seg)
    loc = head_1x1(pa, torch.from_numpy(st["convPb.weight"]), torch.from_numpy(st["convPb.bias"]), mode)
    ids = head_1x1(da, torch.from_numpy(st["convDb.weight"]), torch.from_numpy(st["convDb.bias"]), mode)
    return loc, ids


def ref(x, mode, seg):
    """Materialised upsamplings (the kernel's phase-collapsed 2x2 form has 4 taps per chunk instead of 9: shorter chains)."""
    st = sr
    x = cbr(x, st, "conv1a", 0, mode, seg); x = cbr(x, st, "conv1b", 0, mode, seg); x = cbr(x, st, "conv2a", 0, mode, seg)
    x = cbr(x, st, "conv2b", 0, mode, seg)
    x = F.max_pool2d(x, 2, 2); x = cbr(x, st, "conv3a", 1, mode, seg); x = cbr(x, st, "conv3b", 1, mode, seg)
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = cbr(x, st, "conv4a", 1, mode, seg); x = cbr(x, st, "conv4b", 1, mode, seg); x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = cbr(x, st, "conv5a", 1, mode, seg); x = cbr(x, st, "conv5b", 1, mode, seg); x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = cbr(x, st, "convPa", 1, mode, seg)
    return F.conv2d(x, torch.from_numpy(st["convPb.weight"]), torch.from_numpy(st["convPb.bias"]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--modes", default="rz,rn")
    ap.add_argument("--segs", default="0,9,3", help="MMAs per segment; 'a:b' = a for 64-channel inputs, b for 128-channel inputs")
    ap.add_argument("--patches", type=int, default=12)
    ap.add_argument("--skip-det", action="store_true")
    ap.add_argument("--lo8", type=int, default=None, help="emulate an e4m3 a_lo x w_hi product with this power-of-two shift")
    a = ap.parse_args()
    LO8_SHIFT = a.lo8
    frames = synth.make_frames(a.frames, seed=1)
    with torch.no_grad():
        x = torch.from_numpy(np.stack([oracle.pre_bgr_image(f) for f in frames]))
        loc0, ids0 = oracle.detector_forward(sd, x)
        P = []
        for f in frames:
            r, stg = oracle.pipeline.infer_gray(sd, sr, f, return_stages=True)
            if "patches" in stg:
                P.append(stg["patches"])
        P = torch.from_numpy(np.concatenate(P))[: a.patches, None]
        h0 = oracle.refinenet_forward(sr, P)
        for mode in a.modes.split(","):
            for seg in [({64: int(s.split(":")[0]), 128: int(s.split(":")[1])} if ":" in s else int(s)) for s in a.segs.split(",")]:
                if not a.skip_det:
                    loc, ids = det(x, mode, seg)
                    print(f"mode={mode} seg={seg}: dloc {float((loc - loc0).abs().max()):.3e} dids {float((ids - ids0).abs().max()):.3e} "
                          f"mean signed dloc {float((loc - loc0).mean()):+.2e}", flush=True)
                h = ref(P, mode, seg)
                print(f"mode={mode} seg={seg}: dheat {float((h - h0).abs().max()):.3e} flips "
                      f"{int((h.flatten(1).argmax(1) != h0.flatten(1).argmax(1)).sum())} of {h.shape[0]}", flush=True)
