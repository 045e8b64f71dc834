"""Stage-level parity through the C ABI: each stage of the CUDA path against the oracle / golden fixtures,
fed with the REFERENCE's inputs for that stage so errors cannot compound."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from conftest import split_rows
from deepcharuco_b200 import _native as N

pytestmark = pytest.mark.gpu

# |delta| on logits of magnitude ~1e2.  fp32 CUDA-core path: a different summation order only (measured 2.4e-3).  tcgen05 path: the
# fp16 hi/lo split keeps 22 bits per operand; the tensor core aligns the 16 products of an MMA and the accumulator to the largest
# exponent and truncates each to 26 bits (tools/mma_probe.py reproduces the hardware bit for bit), a one-sided loss that adds up
# over the 36 - 72 MMAs of a tile: measured max 0.056 on these frames (profiles/r2_parity_*.json), 0.0088 with the two-level
# accumulation (DCU_SEG=1, test_strict_accumulation_mode).
LOGIT_TOL = {N.CONV_FFMA: 5e-3, N.CONV_TCGEN05: 0.1}
HEAT_TOL = 1e-5       # |delta| on the 64x64 heat map (range ~[0,1]); measured 5.2e-6 (tcgen05), 1.9e-6 (fp32 path)


def _cuda(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def engine(states):
    e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=16, max_patches=4096)
    yield e
    e.close()


@pytest.mark.parametrize("impl", [N.CONV_FFMA, N.CONV_TCGEN05])
def test_detector_logits(engine, golden_synth, impl):
    g = golden_synth
    engine.set_conv_impl(impl)
    n = g["loc"].shape[0]
    frames = _cuda(g["frames"][:n])
    loc = torch.empty((n, 65, 30, 40), device="cuda")
    ids = torch.empty((n, 17, 30, 40), device="cuda")
    N.check(N.lib().dcu_detector_forward(engine.handle, frames.data_ptr(), n, loc.data_ptr(), ids.data_ptr(), None))
    torch.cuda.synchronize()
    engine.set_conv_impl(N.CONV_DEFAULT)
    dl = (loc.cpu().numpy() - g["loc"])
    di = (ids.cpu().numpy() - g["ids"])
    tol = LOGIT_TOL[impl]
    assert np.abs(dl).max() < tol and np.abs(di).max() < tol, (np.abs(dl).max(), np.abs(di).max())
    # what the decode consumes: the per-cell arg-max of ids, and of loc on every cell the reference keeps
    assert np.array_equal(ids.cpu().numpy().argmax(1), g["ids"].argmax(1))
    kept = (g["loc"].argmax(1) != 64) & (g["ids"].argmax(1) != 16)
    assert np.array_equal(loc.cpu().numpy().argmax(1)[kept], g["loc"].argmax(1)[kept])


def test_detector_f32_entry_equals_u8_entry(engine, golden_synth):
    f = golden_synth["frames"][:2]
    x = _cuda(np.stack([oracle.pre_bgr_image(a)[0] for a in f]))
    fr = _cuda(f)
    out = [torch.empty((2, 65, 30, 40), device="cuda"), torch.empty((2, 17, 30, 40), device="cuda"),
           torch.empty((2, 65, 30, 40), device="cuda"), torch.empty((2, 17, 30, 40), device="cuda")]
    N.check(N.lib().dcu_detector_forward(engine.handle, fr.data_ptr(), 2, out[0].data_ptr(), out[1].data_ptr(), None))
    N.check(N.lib().dcu_detector_forward_f32(engine.handle, x.data_ptr(), 2, out[2].data_ptr(), out[3].data_ptr(), None))
    torch.cuda.synchronize()
    assert torch.equal(out[0], out[2]) and torch.equal(out[1], out[3])     # LUT == numpy normalisation, bit for bit


def _decode(engine, loc, ids, frames, n, patches=True, append=0, total=None):
    counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    offsets = torch.zeros(n, dtype=torch.int32, device="cuda")
    total = torch.zeros(1, dtype=torch.int32, device="cuda") if total is None else total
    kpts = torch.zeros((engine.max_patches, 4), dtype=torch.int32, device="cuda")
    pt = torch.zeros((engine.max_patches, 24, 24), device="cuda") if patches else None
    N.check(N.lib().dcu_decode_gather(engine.handle, loc.data_ptr(), ids.data_ptr(), frames.data_ptr() if frames is not None else None,
                                      n, 16, append, counts.data_ptr(), offsets.data_ptr(), total.data_ptr(), kpts.data_ptr(),
                                      pt.data_ptr() if patches else None, None))
    torch.cuda.synchronize()
    return counts.cpu().numpy(), offsets.cpu().numpy(), int(total.item()), kpts.cpu().numpy(), (pt.cpu().numpy() if patches else None)


def test_decode_gather_bit_exact_on_reference_logits(engine, golden_synth):
    g = golden_synth
    n = g["loc"].shape[0]
    counts, offsets, total, kpts, patches = _decode(engine, _cuda(g["loc"]), _cuda(g["ids"]), _cuda(g["frames"][:n]), n)
    assert counts.tolist() == g["counts"][:n].tolist()
    assert offsets.tolist() == np.concatenate([[0], np.cumsum(counts)[:-1]]).tolist() and total == counts.sum()
    ref_kp = split_rows(g["kpts"], g["counts"])
    ref_id = split_rows(g["ids_found"], g["counts"])
    ref_pt = split_rows(g["patches"], g["counts"][:n])
    for f in range(n):
        rows = kpts[offsets[f]:offsets[f] + counts[f]]
        # engine order = (id, cell); reference order = row-major: compare as the reference's stable sort by id
        order = sorted(range(len(ref_id[f])), key=lambda i: ref_id[f][i])
        assert np.array_equal(rows[:, :2], ref_kp[f][order]) and np.array_equal(rows[:, 2], ref_id[f][order])
        # sorting the engine rows by cell restores pred_to_keypoints order
        back = rows[np.argsort(rows[:, 3], kind="stable")]
        assert np.array_equal(back[:, :2], ref_kp[f]) and np.array_equal(back[:, 2], ref_id[f])
        assert np.array_equal(patches[offsets[f]:offsets[f] + counts[f]], ref_pt[f][order])     # fp32 patches bit-exact


def test_decode_edge_cases(engine):
    """ties -> first max; loc dustbin 64; ids dustbin; all-dustbin frame -> K = 0; append mode continues numbering."""
    loc = np.full((3, 65, 30, 40), -1.0, np.float32)
    ids = np.full((3, 17, 30, 40), -1.0, np.float32)
    loc[:, 64] = 0.0                                   # everything dustbin by default
    loc[0, 5, 2, 3] = loc[0, 9, 2, 3] = 3.0            # tie -> 5 -> x = 8*3+5, y = 8*2+0
    ids[0, 2, 2, 3] = ids[0, 7, 2, 3] = 1.0            # tie -> id 2
    loc[0, 63, 29, 39] = 1.0; ids[0, 0, 29, 39] = 1.0  # last cell, last pixel, id 0
    loc[0, 10, 0, 0] = 1.0; ids[0, 16, 0, 0] = 5.0     # ids dustbin wins -> dropped
    loc[2, 0, 7, 7] = 2.0; ids[2, 15, 7, 7] = 2.0
    frames = np.zeros((3, 240, 320), np.uint8)
    counts, offsets, total, kpts, _ = _decode(engine, _cuda(loc), _cuda(ids), _cuda(frames), 3)
    assert counts.tolist() == [2, 0, 1] and offsets.tolist() == [0, 2, 2] and total == 3
    assert kpts[0].tolist() == [319, 239, 0, 29 * 40 + 39]
    assert kpts[1].tolist() == [29, 16, 2, 2 * 40 + 3]
    assert kpts[2].tolist() == [56, 56, 15, 7 * 40 + 7]
    want_k, want_i = oracle.pred_to_keypoints(loc[:1], ids[:1], 16)
    assert sorted(map(tuple, kpts[:2, :2].tolist())) == sorted(map(tuple, want_k.tolist()))
    # append: a second call continues after `total`
    tot = torch.tensor([3], dtype=torch.int32, device="cuda")
    c2, o2, t2, k2, _ = _decode(engine, _cuda(loc), _cuda(ids), _cuda(frames), 3, append=1, total=tot)
    assert o2.tolist() == [3, 5, 5] and t2 == 6


def test_patch_zero_padding_at_borders(engine):
    """A corner at (0,0) / (319,239): the window hangs outside the frame and must read 0.0 (= gray 128 normalised)."""
    loc = np.full((1, 65, 30, 40), -1.0, np.float32); loc[:, 64] = 0.0
    ids = np.full((1, 17, 30, 40), -1.0, np.float32)
    loc[0, 0, 0, 0] = 1.0; ids[0, 1, 0, 0] = 1.0
    loc[0, 63, 29, 39] = 1.0; ids[0, 2, 29, 39] = 1.0
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (1, 240, 320)).astype(np.uint8)
    counts, offsets, total, kpts, patches = _decode(engine, _cuda(loc), _cuda(ids), _cuda(frame), 1)
    img = oracle.pre_bgr_image(frame[0])
    want = oracle.extract_patches(img, kpts[:2, :2].astype(np.int64))
    assert np.array_equal(patches[:2], want)
    assert (want[0][:12, :] == 0).all() and (want[1][13:, :] == 0).all()


def test_extract_patches_entry(engine, golden_synth):
    g = golden_synth
    img = oracle.pre_bgr_image(g["frames"][0])
    kp = split_rows(g["kpts"], g["counts"])[0]
    out = torch.empty((len(kp), 24, 24), device="cuda")
    d_img, d_kp = _cuda(img), _cuda(kp.astype(np.int32))       # keep the device tensors alive across the async call
    N.check(N.lib().dcu_extract_patches(engine.handle, d_img.data_ptr(), d_kp.data_ptr(), len(kp), out.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), oracle.extract_patches(img, kp))


def test_refinenet_on_reference_patches(engine, golden_synth):
    g = golden_synth
    p = g["patches"].shape[0]
    kp = g["kpts"][:p].astype(np.int32)
    corners = torch.empty((p, 2), dtype=torch.int32, device="cuda")
    refined = torch.empty((p, 2), device="cuda")
    heat = torch.empty((p, 64, 64), device="cuda")
    d_patches, d_kp = _cuda(g["patches"]), _cuda(kp)            # keep alive: the call is asynchronous
    N.check(N.lib().dcu_refine_forward(engine.handle, d_patches.data_ptr(), d_kp.data_ptr(), 2, p,
                                       corners.data_ptr(), refined.data_ptr(), heat.data_ptr(), None))
    torch.cuda.synchronize()
    dh = np.abs(heat.cpu().numpy() - g["heat"]).max()
    assert dh < HEAT_TOL, dh
    got = corners.cpu().numpy()
    want = g["corners"][:p]
    assert np.array_equal(got, want)                    # every 64x64 arg-max of the golden set equals the reference's
    ref_refined = (want.astype(np.float32) - 32) / 8 + kp.astype(np.float32)
    assert np.array_equal(refined.cpu().numpy(), ref_refined)
    # the engine's own arg-max is consistent with its own heat map (first max wins)
    assert np.array_equal(got, oracle.bargmax2d(heat.cpu().numpy()))


def test_refinenet_16384_patches_periodic(states, golden_synth):
    """BASELINE config 4 size (16384 patches, four chunks of 4096): the golden patches cycled.  A patch's result may not depend on
    its position in the batch (flat pixel runs put patches at every offset inside the 128-pixel MMA tiles), so the output is
    periodic and equal to the small-batch result; refined = kp + (corner - 32) / 8 exactly."""
    g = golden_synth
    p0 = g["patches"].shape[0]
    e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=16, max_patches=16384)
    try:
        reps = (16384 + p0 - 1) // p0
        patches = _cuda(np.tile(g["patches"], (reps, 1, 1))[:16384])
        kp = _cuda(np.tile(g["kpts"][:p0].astype(np.int32), (reps, 1))[:16384])
        corners = torch.empty((16384, 2), dtype=torch.int32, device="cuda")
        refined = torch.empty((16384, 2), device="cuda")
        N.check(N.lib().dcu_refine_forward(e.handle, patches.data_ptr(), kp.data_ptr(), 2, 16384, corners.data_ptr(),
                                           refined.data_ptr(), None, None))
        small_c = torch.empty((p0, 2), dtype=torch.int32, device="cuda")
        small_r = torch.empty((p0, 2), device="cuda")
        N.check(N.lib().dcu_refine_forward(e.handle, patches.data_ptr(), kp.data_ptr(), 2, p0, small_c.data_ptr(),
                                           small_r.data_ptr(), None, None))
        torch.cuda.synchronize()
        c, r = corners.cpu().numpy(), refined.cpu().numpy()
        idx = np.arange(16384) % p0
        assert np.array_equal(c, small_c.cpu().numpy()[idx]) and np.array_equal(r, small_r.cpu().numpy()[idx])
        assert np.array_equal(r, (c.astype(np.float32) - 32) / 8 + kp.cpu().numpy().astype(np.float32))
        assert np.array_equal(c, g["corners"][:p0][idx])                    # and equal to the reference's arg-maxes
    finally:
        e.close()


def _layer(engine, net, layer, impl, x, out_shape):
    n, c, h, w = x.shape
    xin = _cuda(x)
    out = torch.full(tuple(out_shape), float("nan"), device="cuda")
    N.check(N.lib().dcu_debug_conv_layer(engine.handle, net, layer, impl, xin.data_ptr(), n, h, w, out.data_ptr(), None))
    torch.cuda.synchronize()      # (the entry point synchronises too; xin must outlive the kernels)
    return out.cpu().numpy()


@pytest.mark.parametrize("impl", [N.CONV_FFMA, N.CONV_TCGEN05])
def test_every_conv_layer_against_oracle(engine, states, golden_synth, impl):
    """Feed each 3x3 layer the ORACLE's input activation and compare its output with the oracle's."""
    g = golden_synth
    sd, sr = states
    x0 = torch.from_numpy(np.stack([oracle.pre_bgr_image(f) for f in g["frames"][:2]]))
    _, _, fd = oracle.detector_forward(sd, x0, return_features=True)
    det_chain = [("conv1a", x0), ("conv1b", fd["conv1a"]), ("conv2a", fd["conv1b"]), ("conv2b", fd["conv2a"]),
                 ("conv3a", fd["conv2b"]), ("conv3b", fd["conv3a"]), ("conv4a", fd["conv3b"]), ("conv4b", fd["conv4a"])]
    for li, (name, xin) in enumerate(det_chain):
        want = fd[name].numpy()
        got = _layer(engine, 0, li, N.CONV_FFMA if li == 0 else impl, xin.numpy(), want.shape)
        err = np.abs(got - want).max()
        assert err < 1e-4 * max(1.0, np.abs(want).max()), (name, err)
    want = torch.cat([fd["convPa"], fd["convDa"]], 1).numpy()
    got = _layer(engine, 0, 8, impl, fd["conv4b"].numpy(), want.shape)
    assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())

    p0 = torch.from_numpy(g["patches"][:6])[:, None]
    _, fr = oracle.refinenet_forward(sr, p0, return_features=True)
    names = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "conv5a", "conv5b", "convPa"]
    xin = p0
    for li, name in enumerate(names):
        want = fr[name].numpy()
        got = _layer(engine, 1, li, N.CONV_FFMA if li == 0 else impl, xin.numpy(), want.shape)
        err = np.abs(got - want).max()
        assert err < 1e-4 * max(1.0, np.abs(want).max()), (name, err)
        xin = fr[name]


def test_strict_accumulation_mode(states, golden_synth, monkeypatch):
    """DCU_SEG=1: two-level accumulation (one 16-channel chunk per chain in tensor memory, fp32 round-to-nearest adds in registers):
    the logits move ~6x closer to the oracle (max |delta| 0.0088 instead of 0.056 on the parity set; here 3 golden frames)."""
    g = golden_synth
    n = g["loc"].shape[0]
    errs = {}
    for seg in ("0", "1"):
        monkeypatch.setenv("DCU_SEG", seg)
        e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=8, max_patches=1024)
        try:
            frames = _cuda(g["frames"][:n])
            loc = torch.empty((n, 65, 30, 40), device="cuda"); ids = torch.empty((n, 17, 30, 40), device="cuda")
            N.check(N.lib().dcu_detector_forward(e.handle, frames.data_ptr(), n, loc.data_ptr(), ids.data_ptr(), None))
            p = g["patches"].shape[0]
            kp = _cuda(g["kpts"][:p].astype(np.int32)); pt = _cuda(g["patches"])
            corners = torch.empty((p, 2), dtype=torch.int32, device="cuda"); refined = torch.empty((p, 2), device="cuda")
            heat = torch.empty((p, 64, 64), device="cuda")
            N.check(N.lib().dcu_refine_forward(e.handle, pt.data_ptr(), kp.data_ptr(), 2, p, corners.data_ptr(), refined.data_ptr(), heat.data_ptr(), None))
            torch.cuda.synchronize()
            errs[seg] = (float(np.abs(loc.cpu().numpy() - g["loc"]).max()), float(np.abs(heat.cpu().numpy() - g["heat"]).max()))
            assert np.array_equal(corners.cpu().numpy(), g["corners"][:p])
        finally:
            e.close()
    print("max |dloc|, |dheat|: whole-tile chains", errs["0"], " two-level", errs["1"])
    assert errs["1"][0] < 0.02 and errs["1"][1] < 2.5e-6
    assert errs["1"][0] < 0.6 * errs["0"][0] and errs["1"][1] < 0.6 * errs["0"][1]
