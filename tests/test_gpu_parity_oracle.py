"""CUDA engine vs. the oracle run on this box's CPU, on seeded synthetic frames that are NOT in the golden set,
plus size-independent properties at the BASELINE batch sizes."""
import numpy as np
import pytest

import deepcharuco_b200 as dc
from deepcharuco_b200 import synth, _native as N
import parity

pytestmark = pytest.mark.gpu


def test_parity_48_frames_seed1(models, states):
    deepc, refinenet = models
    frames = synth.make_frames(48, 240, 320, seed=1)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    reps = [parity.compare_frame(states, f, r, w) for f, r, w in zip(frames, refined, raw)]
    tot = parity.summarise(reps)
    print("PARITY 320x240 seed1:", tot)
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)


def test_parity_640x480(models, states):
    deepc, refinenet = models
    frames = synth.make_frames(4, 480, 640, seed=5)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    tot = parity.summarise([parity.compare_frame(states, f, r, w) for f, r, w in zip(frames, refined, raw)])
    print("PARITY 640x480 seed5:", tot)
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)


@pytest.mark.parametrize("hw", [(200, 296), (24, 24), (136, 72)])
def test_parity_awkward_sizes(models, states, hw):
    """Frame sizes that are multiples of 8 but of nothing else the kernels tile by: odd cell grids (25 x 37, 3 x 3, 17 x 9), maps whose
    pooled sizes are odd, tiles that overhang on every edge; 24 x 24 is the smallest frame the engine accepts."""
    deepc, refinenet = models
    H, W = hw
    big = synth.make_frames(6, 240, 320, seed=9)
    frames = np.ascontiguousarray(big[:, 20:20 + H, 12:12 + W])
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    tot = parity.summarise([parity.compare_frame(states, f, r, w) for f, r, w in zip(frames, refined, raw)])
    print("PARITY %dx%d:" % (W, H), tot)
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)
    # small crops rarely keep a corner, so also compare what the detector itself produces at this size
    import torch
    import oracle
    eng = deepc._ctx.engine(H, W, max_batch=len(frames))
    fr = torch.from_numpy(frames).cuda()
    loc = torch.empty((len(frames), 65, H // 8, W // 8), device="cuda")
    ids = torch.empty((len(frames), 17, H // 8, W // 8), device="cuda")
    N.check(N.lib().dcu_detector_forward(eng.handle, fr.data_ptr(), len(frames), loc.data_ptr(), ids.data_ptr(), None))
    torch.cuda.synchronize()
    x = torch.from_numpy(np.stack([oracle.pre_bgr_image(f) for f in frames]))
    wl, wi = oracle.detector_forward(states[0], x)
    wl, wi = wl.numpy(), wi.numpy()
    assert np.abs(loc.cpu().numpy() - wl).max() < 0.1 and np.abs(ids.cpu().numpy() - wi).max() < 0.1     # logits are O(100)
    assert np.array_equal(ids.cpu().numpy().argmax(1), wi.argmax(1))


def test_properties_at_batch_256(models):
    """BASELINE config 3 size.  Frames are independent, so (a) a permuted batch gives permuted results,
    (b) repeating the call is bit-identical (idempotent / deterministic), (c) duplicates give identical rows,
    (d) every refined corner stays within the sub-pixel window of its raw pixel: kp + [-4, 3.875], step 1/8."""
    deepc, refinenet = models
    pool = synth.make_frames(32, 240, 320, seed=2)
    frames = synth.tile_frames(pool, 256)
    a = dc.infer_batch(frames, 16, deepc, refinenet)
    b = dc.infer_batch(frames, 16, deepc, refinenet)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    perm = np.random.default_rng(0).permutation(256)
    c = dc.infer_batch(frames[perm], 16, deepc, refinenet)
    assert all(np.array_equal(c[i], a[perm[i]]) for i in range(256))
    assert all(np.array_equal(a[i], a[i + 32]) for i in range(0, 224))
    raw = dc.infer_batch(frames, 16, deepc, None)
    for r, w in zip(a, raw):
        if r.size == 0:
            assert w.size == 0
            continue
        assert np.array_equal(r[:, 2], w[:, 2]) and np.all(np.diff(r[:, 2]) >= 0)       # sorted by id
        d = r[:, :2] - w[:, :2]
        assert d.min() >= -4.0 and d.max() <= 3.875 and np.array_equal(d * 8, np.round(d * 8))
