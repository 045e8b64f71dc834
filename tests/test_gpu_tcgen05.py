"""The tcgen05 (fp16 hi/lo split, 3 products) convolution path: per layer against the fp32 CUDA-core kernel and the oracle, then the
whole pipeline against the golden fixtures and the oracle.  Runs last (a kernel fault poisons the CUDA context)."""
import numpy as np
import pytest
import torch

import deepcharuco_b200 as dc
import oracle
import parity
from conftest import split_rows
from deepcharuco_b200 import _native as N, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(states):
    e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=16, max_patches=4096)
    yield e
    e.close()


@pytest.fixture(scope="module")
def tc_models():
    deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda")
    deepc._ctx.set_conv_impl(N.CONV_TCGEN05)
    return deepc, refinenet


SHAPES = [  # net, layer, cin, h, w, cout, out_h, out_w
    (0, 1, 64, 240, 320, 64, 120, 160), (0, 2, 64, 120, 160, 64, 120, 160), (0, 4, 64, 60, 80, 128, 60, 80),
    (0, 5, 128, 60, 80, 128, 30, 40), (0, 6, 128, 30, 40, 128, 30, 40), (0, 8, 128, 30, 40, 512, 30, 40),
    (1, 1, 64, 22, 22, 64, 20, 20), (1, 2, 64, 20, 20, 128, 18, 18), (1, 3, 128, 18, 18, 128, 8, 8),
    (1, 4, 128, 8, 8, 128, 8, 8), (1, 5, 128, 8, 8, 128, 16, 16), (1, 6, 128, 16, 16, 128, 16, 16),
    (1, 7, 128, 16, 16, 128, 32, 32), (1, 8, 128, 32, 32, 64, 32, 32), (1, 9, 64, 32, 32, 64, 64, 64), (1, 10, 64, 64, 64, 64, 64, 64),
]
# RefineNet layers whose input is always a 2x nearest upsampling (refinenet.py:66,71,76): the tcgen05 path folds the upsampling
# into the convolution (2x2 phase kernels on the low-resolution tensor), so they are tested on upsampled inputs
UPSAMPLED_INPUT = {(1, 6), (1, 8), (1, 10)}


# n = 37 patches: RefineNet's small maps run as one flat run of pixels (conv_tc2.cu FLAT mode), so images straddle tile and
# CTA-pair boundaries and the last pair is ragged
@pytest.mark.parametrize("shape,n", [(s, 3) for s in SHAPES] + [(s, 37) for s in SHAPES if s[0] == 1 and s[1] <= 6])
def test_layer_tcgen05_matches_fp32_kernel(engine, shape, n):
    net, layer, cin, h, w, cout, oh, ow = shape
    rng = np.random.default_rng(layer * 7 + net + n)
    if (net, layer) in UPSAMPLED_INPUT:
        x = np.maximum(rng.standard_normal((n, cin, h // 2, w // 2)).astype(np.float32), 0).repeat(2, axis=2).repeat(2, axis=3)
        x = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    else:
        x = torch.from_numpy(np.maximum(rng.standard_normal((n, cin, h, w)).astype(np.float32), 0)).cuda()
    outs = []
    for impl in (N.CONV_FFMA, N.CONV_TCGEN05):
        out = torch.full((n, cout, oh, ow), float("nan"), device="cuda")
        N.check(N.lib().dcu_debug_conv_layer(engine.handle, net, layer, impl, x.data_ptr(), n, h, w, out.data_ptr(), None))
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    a, b = outs
    assert not np.isnan(b).any()
    # the fp16 hi/lo split keeps ~22 bits per operand: agreement with the fp32 FMA kernel at fp32-noise level
    assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(a).max()), np.abs(a - b).max()


def test_pipeline_golden_tcgen05(tc_models, golden_sample, golden_synth):
    deepc, refinenet = tc_models
    kp, _ = dc.infer_image(golden_sample["bgr"], 16, deepc, refinenet)
    assert np.array_equal(kp[:, 2], golden_sample["out_refined"][:, 2])
    assert np.abs(kp[:, :2] - golden_sample["out_refined"][:, :2]).max() <= 1e-3
    raw, _ = dc.infer_image(golden_sample["bgr"], 16, deepc, None)
    assert np.array_equal(raw, golden_sample["out_raw"])
    g = golden_synth
    res = dc.infer_batch(g["frames"], 16, deepc, refinenet)
    for got, w in zip(res, split_rows(g["out_refined"], g["counts"])):
        assert got.shape == w.shape and np.array_equal(got[:, 2], w[:, 2])
        assert np.abs(got[:, :2] - w[:, :2]).max() <= 1e-3


def test_parity_tcgen05_vs_oracle(tc_models, states):
    deepc, refinenet = tc_models
    frames = synth.make_frames(48, 240, 320, seed=1)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    tot = parity.summarise([parity.compare_frame(states, f, r, w) for f, r, w in zip(frames, refined, raw)])
    print("PARITY tcgen05 320x240 seed1:", tot)
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)


def test_tcgen05_deterministic_and_batch_invariant(tc_models):
    deepc, refinenet = tc_models
    frames = synth.tile_frames(synth.make_frames(16, 240, 320, seed=2), 64)
    a = dc.infer_batch(frames, 16, deepc, refinenet)
    b = dc.infer_batch(frames, 16, deepc, refinenet)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    one = dc.infer_batch(frames[5:6], 16, deepc, refinenet)[0]
    assert np.array_equal(one, a[5])


def test_fused_first_layer_is_bit_identical(states, golden_synth, monkeypatch):
    """conv1a computed inside conv1b's kernel (conv_tc2.cu FIRST mode, DCU_FUSE_FIRST=1; off by default) vs the separate conv1a kernel + HBM round trip
    (DCU_FUSE_FIRST=0): same FMA order for conv1a, same MMAs for conv1b -> bit-identical logits, for u8 and fp32 inputs, at a
    size with ragged tiles too."""
    import torch
    monkeypatch.setenv("DCU_SEG", "0")       # the fused kernel keeps whole-tile accumulation chains (no two-level accumulation there)
    for (H, W) in ((240, 320), (200, 296)):
        frames = np.ascontiguousarray(golden_synth["frames"][:5, :H, :W])
        outs = {}
        for fuse in ("1", "0"):
            monkeypatch.setenv("DCU_FUSE_FIRST", fuse)
            e = N.Engine(states[0], states[1], H, W, 16, 0, max_batch=8, max_patches=1024)
            try:
                fr = torch.from_numpy(frames).cuda()
                x = torch.from_numpy(np.stack([oracle.pre_bgr_image(f)[0] for f in frames])).cuda()
                res = []
                for entry, src in ((N.lib().dcu_detector_forward, fr), (N.lib().dcu_detector_forward_f32, x)):
                    loc = torch.empty((5, 65, H // 8, W // 8), device="cuda")
                    ids = torch.empty((5, 17, H // 8, W // 8), device="cuda")
                    N.check(entry(e.handle, src.data_ptr(), 5, loc.data_ptr(), ids.data_ptr(), None))
                    torch.cuda.synchronize()
                    res += [loc.cpu().numpy(), ids.cpu().numpy()]
                outs[fuse] = res
            finally:
                e.close()
        for a, b in zip(outs["1"], outs["0"]):
            assert np.array_equal(a, b)
        assert np.array_equal(outs["1"][0], outs["1"][2])          # u8 entry == fp32 entry


def test_argmax_heads_equal_logit_decode(states, monkeypatch):
    """Fused pipeline with the per-cell arg-max taken in the 1x1 head epilogues (default) vs heads that write fp32 logits which the
    decode kernel re-reads (DCU_ARG_HEADS=0): the arg-max is taken on the same values, so every result row is identical."""
    frames = synth.make_frames(40, 240, 320, seed=17)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DCU_ARG_HEADS", mode)
        e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=40, max_patches=4096)
        try:
            outs[mode] = [a.copy() if a is not None else None for a in e.infer_batch_host(frames, 16, True)]
        finally:
            e.close()
    for a, b in zip(outs["1"], outs["0"]):
        assert np.array_equal(a, b)
    assert outs["1"][0].sum() > 100          # the frames do contain corners
