"""What does the B200 tensor core do when tcgen05.mma (kind::f16, fp32 accumulate) adds into a running accumulator?

Run ON THE GPU BOX.  Uses the production kernel (conv_tc2_kernel through dcu_debug_conv_layer) with a crafted layer:
inputs and scaled weights are exactly representable in fp16 (so a_lo = w_lo = 0 and the result is ONE chain of exact
fp16 x fp16 products), spatially constant images (every interior pixel sees the same 576 / 1152 products), BatchNorm =
identity.  Channel c + C/2 carries the same input as channel c with the weight -w + delta, so the accumulator climbs to a
few hundred during the first half of the chain and returns to a small, exactly representable value: whatever the
accumulating additions lost on the way is then visible EXACTLY in the output.

Each sample (image, output channel) is compared with CPU models of the accumulate step (vectorised, exact in float64):
  exact   : no loss at all
  rn / rz / rd : the 16 products of one MMA are summed exactly and added to the accumulator with round-to-nearest / toward zero /
                 toward -inf to fp32
  align(g, z|d): accumulator and the 16 products are aligned to the largest exponent and EACH is truncated (toward zero / toward
                 -inf) to 24+g bits there before the sum (which is then truncated the same way to fp32)
and the tool prints how many samples each model reproduces bit for bit, plus the signed error in units of the quantum of the
largest partial sum.

    python tools/mma_probe.py [--out gpurun_out/mma_probe.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from deepcharuco_b200 import _native as N, weights_io as W  # noqa: E402


def identity_bn(state, bn, c):
    state[bn + ".weight"] = np.ones(c, np.float32)
    state[bn + ".bias"] = np.zeros(c, np.float32)
    state[bn + ".running_mean"] = np.zeros(c, np.float32)
    for v in (np.float32(1.0) - N.BN_EPS, np.float32(1.0), np.float32(1.0) - np.float32(2) * N.BN_EPS):
        state[bn + ".running_var"] = np.full(c, v, np.float32)
        a, b = N.fold_bn(state, "conv" + bn[2:])
        if np.all(a == 1.0) and np.all(b == 0.0):
            return
    raise RuntimeError("no running_var gives alpha == 1 exactly")


def rn32(x):
    return x.astype(np.float32).astype(np.float64)


def rz32(x):
    y = x.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x)
    y = np.where(over, np.nextafter(y, np.float32(0)), y)
    return y.astype(np.float64)


def rd32(x):
    y = x.astype(np.float32)
    over = y.astype(np.float64) > x
    y = np.where(over, np.nextafter(y, np.float32(-np.inf)), y)
    return y.astype(np.float64)


def expo(x):
    """floor(log2|x|), very negative for 0."""
    m, e = np.frexp(x)
    return np.where(x == 0, -1000, e - 1)


def model_chain(acc0, prods, kind, g=0):
    """prods: [steps][16][samples] exact products (float64).  Returns the final accumulator per sample."""
    acc = acc0.copy()
    for p in prods:
        if kind == "exact":
            acc = acc + p.sum(0)
        elif kind == "rn":
            acc = rn32(acc + p.sum(0))
        elif kind == "rz":
            acc = rz32(acc + p.sum(0))
        elif kind == "rd":
            acc = rd32(acc + p.sum(0))
        else:   # align_z / align_d
            terms = np.concatenate([acc[None], p], 0)
            E = expo(terms).max(0)
            q = np.exp2((E - 23 - g).astype(np.float64))
            t = terms / q
            t = np.trunc(t) if kind == "align_z" else np.floor(t)
            s = t.sum(0) * q
            acc = rz32(s) if kind == "align_z" else rd32(s)
    return acc


def run_case(name, conv, bn, layer, cin, cout, sign, rng, n_img=64, hw=16, fake_hw=None):
    sd, sr = W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)
    sd = dict(sd)
    half = cin // 2
    # weights m * 2^-10, |m| <= 1023 (so w * 2^15 is an exact fp16); second half = -first half + delta
    m = rng.integers(200, 1023, (cout, half, 3, 3)).astype(np.float64)
    if sign == "mixed":
        m *= rng.choice([-1.0, 1.0], m.shape)
    elif sign == "neg":
        m = -m            # the accumulator runs negative first: tells truncation toward zero from truncation toward -inf
    delta = np.zeros_like(m)
    delta[:, :, 1, 1] = 1.0                                      # + 2^-10 per channel at the centre tap: the final value is positive (ReLU)
    w = np.concatenate([m, -m + delta], 1) * 2.0 ** -10
    sd[conv + ".weight"] = w.astype(np.float32)
    sd[conv + ".bias"] = np.zeros(cout, np.float32)
    identity_bn(sd, bn, cout)
    a = 1.0 + rng.integers(0, 1024, (n_img, half)).astype(np.float64) / 1024.0        # exact fp16 in [1, 2)
    a = np.concatenate([a, a], 1)
    x = np.broadcast_to(a[:, :, None, None], (n_img, cin, hw, hw)).astype(np.float32).copy()
    if fake_hw is not None:
        got = ffma = None
    else:
      eng = N.Engine(sd, sr, 240, 320, 16, 0, max_batch=8, max_patches=256)
      try:
        xin = torch.from_numpy(x).cuda()
        out = torch.full((n_img, cout, hw, hw), float("nan"), device="cuda")
        N.check(N.lib().dcu_debug_conv_layer(eng.handle, 0, layer, N.CONV_TCGEN05, xin.data_ptr(), n_img, hw, hw, out.data_ptr(), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy()[:, :, hw // 2, hw // 2].astype(np.float64)           # [img][cout], interior pixel
        out2 = torch.full((n_img, cout, hw, hw), float("nan"), device="cuda")
        N.check(N.lib().dcu_debug_conv_layer(eng.handle, 0, layer, N.CONV_FFMA, xin.data_ptr(), n_img, hw, hw, out2.data_ptr(), None))
        torch.cuda.synchronize()
        ffma = out2.cpu().numpy()[:, :, hw // 2, hw // 2].astype(np.float64)
      finally:
        eng.close()
    # the kernel's MMA order: chunk q (16 channels) -> ky -> kx; products scaled by 2^15 inside, undone exactly afterwards
    wd = sd[conv + ".weight"].astype(np.float64)
    prods = []
    for q in range(cin // 16):
        for ky in range(3):
            for kx in range(3):
                # [16][img][cout]
                prods.append(np.stack([a[:, q * 16 + j][:, None] * wd[:, q * 16 + j, ky, kx][None, :] for j in range(16)], 0).reshape(16, -1))
    exact = model_chain(np.zeros(n_img * cout), prods, "exact")
    peak = np.abs(np.cumsum(np.stack([p.sum(0) for p in prods], 0), 0)).max(0)
    quantum = np.exp2(expo(peak) - 23.0)
    res = dict(name=name, layer=conv, chain=len(prods), samples=int(exact.size), sign=sign,
               peak_partial_sum=float(peak.mean()), final_mean=float(exact.mean()))
    if fake_hw is not None:          # CPU self-test of the analysis: pretend the hardware follows one of the models
        got = model_chain(np.zeros(n_img * cout), prods, fake_hw[0], fake_hw[1])
        ffma = got
    g_flat = got.reshape(-1)
    # ReLU clamps negatives: keep samples whose exact value is comfortably positive
    ok = exact > 64 * quantum
    res["used"] = int(ok.sum())
    err = (g_flat - exact)[ok] / quantum[ok]
    res["hw_err_quanta"] = dict(mean=float(err.mean()), min=float(err.min()), max=float(err.max()), std=float(err.std()))
    errf = (ffma.reshape(-1) - exact)[ok] / quantum[ok]
    res["ffma_err_quanta"] = dict(mean=float(errf.mean()), min=float(errf.min()), max=float(errf.max()), std=float(errf.std()))
    models = [("exact", 0), ("rn", 0), ("rz", 0), ("rd", 0)] + [(k, g) for k in ("align_z", "align_d") for g in (0, 1, 2, 3, 4)]
    res["models"] = {}
    for kind, g in models:
        pred = model_chain(np.zeros(n_img * cout), prods, kind, g)
        match = int((pred[ok] == g_flat[ok]).sum())
        d = (pred - exact)[ok] / quantum[ok]
        res["models"][f"{kind}{g if kind.startswith('align') else ''}"] = dict(
            bit_exact_matches=match, model_err_mean=float(d.mean()), resid_mean=float(((g_flat - pred)[ok] / quantum[ok]).mean()),
            resid_absmax=float(np.abs((g_flat - pred)[ok] / quantum[ok]).max()))
    best = max(res["models"].items(), key=lambda kv: kv[1]["bit_exact_matches"])
    res["best_model"] = best[0]
    print(f"{name}: chain {res['chain']} MMAs, {res['used']} samples, HW error {res['hw_err_quanta']} quanta of the peak partial sum; "
          f"best model {best[0]} reproduces {best[1]['bit_exact_matches']}/{res['used']} bit for bit", flush=True)
    for k, v in res["models"].items():
        print(f"    {k:10s} matches {v['bit_exact_matches']:5d}  model mean err {v['model_err_mean']:+8.3f}  residual mean {v['resid_mean']:+8.3f} max {v['resid_absmax']:.3f}")
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/mma_probe.json")
    ap.add_argument("--self-test", action="store_true", help="no GPU: check that the analysis identifies a known model")
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    report = []
    if a.self_test:
        for fake in (("rz", 0), ("align_d", 2)):
            r = run_case(f"self-test {fake}", "conv2a", "bn2a", 2, 64, 64, "mixed", rng, n_img=8, fake_hw=fake)
            want = fake[0] + (str(fake[1]) if fake[0].startswith("align") else "")
            assert r["models"][want]["bit_exact_matches"] == r["used"], r
        print("self-test ok")
        sys.exit(0)
    for sign in ("pos", "mixed", "neg"):
        report.append(run_case(f"conv2a 64->64 ({sign} weights)", "conv2a", "bn2a", 2, 64, 64, sign, rng))
    report.append(run_case("conv4a 128->128 (pos weights)", "conv4a", "bn4a", 6, 128, 128, "pos", rng))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)
