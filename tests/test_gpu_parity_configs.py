"""Oracle parity at the BASELINE configurations themselves (BASELINE.json configs 3 and 5, plus the paths only they reach),
through the public surface / the C ABI at the engine's production settings.

  * config 3: 256 x 320x240 frames in one call -- four full-resolution micro-batches (mb1 = 64);
  * config 5 shard: 256 x 640x480 frames in one `dcu_infer_batch_host` call -- mb2 = 64 at this size, so the engine loops over four
    groups and the decode of groups 1..3 APPENDS to the rows of the groups before (engine.cu: dcu_infer_batch);
  * the same append path at 320x240 with DCU_MB2=64 and a ragged last group;
  * detectors with n_ids != 16.

Gate (tests/parity.py): kept cells, ids and raw pixels bit-exact; a refined corner may differ only where the oracle's own
heat map is tied within 1e-5 -- and the number of such flips is pinned, not bounded.
"""
import numpy as np
import pytest

import deepcharuco_b200 as dc
import oracle
import parity
from deepcharuco_b200 import _native as N, synth

pytestmark = pytest.mark.gpu


def _compare(states, frames, refined, raw):
    return parity.summarise([parity.compare_frame(states, f, r, w) for f, r, w in zip(frames, refined, raw)])


def test_parity_config3_256_frames(models, states):
    """BASELINE config 3: batch = 256, 320x240, full pipeline, against the oracle frame by frame (3675 corners)."""
    deepc, refinenet = models
    frames = synth.make_frames(256, 240, 320, seed=1)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    tot = _compare(states, frames, refined, raw)
    print("PARITY config 3 (256 x 320x240, seed 1):", tot)
    assert tot["K"] > 3000
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)


def test_parity_config5_shard_256x640x480(models, states):
    """BASELINE config 5, one GPU's shard: 256 frames of 640x480 in ONE host call = four groups of mb2 = 64 frames, groups 1..3 in
    append mode.  64 distinct frames (4 boards each); every group sees them in a different rotation, so a wrong row offset or a
    group that overwrites its predecessor cannot cancel out.  The oracle runs once per distinct frame."""
    deepc, refinenet = models
    pool = synth.make_frames(64, 480, 640, seed=7)
    order = np.concatenate([np.roll(np.arange(64), 5 * g) for g in range(4)])
    frames = np.ascontiguousarray(pool[order])
    eng = deepc._ctx.engine(480, 640, max_batch=256)
    refined = dc.infer_batch(frames, 16, deepc, refinenet)
    raw = dc.infer_batch(frames, 16, deepc, None)
    assert deepc._ctx.engine(480, 640, max_batch=256).max_batch >= 256
    cache = {}
    reps = [parity.compare_frame(states, frames[i], refined[i], raw[i], cache=cache, key=int(order[i])) for i in range(256)]
    tot = parity.summarise(reps)
    print("PARITY config 5 shard (256 x 640x480, seed 7):", tot)
    assert tot["K"] > 8000
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)
    del eng


@pytest.mark.parametrize("use_ref", [True, False])
def test_parity_append_path_small_groups(states, monkeypatch, use_ref):
    """The multi-group path at 320x240: DCU_MB2=64 / DCU_MB1=32 and 200 frames = groups of 64, 64, 64 and a ragged 8."""
    monkeypatch.setenv("DCU_MB1", "32")
    monkeypatch.setenv("DCU_MB2", "64")
    frames = synth.make_frames(200, 240, 320, seed=11)
    e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=200, max_patches=200 * 64)
    try:
        counts, offsets, kpts, refined = e.infer_batch_host(frames, 16, use_ref)
    finally:
        e.close()
    assert offsets.tolist() == np.concatenate([[0], np.cumsum(counts)[:-1]]).tolist()
    from deepcharuco_b200.inference import _rows_to_frames
    got = _rows_to_frames(counts, offsets, kpts, refined)
    if use_ref:
        tot = parity.summarise([parity.compare_frame(states, f, g) for f, g in zip(frames, got)])
        print("PARITY append path (200 x 320x240, groups of 64):", tot)
        assert tot["K"] > 2000
        parity.assert_parity(tot)
        parity.assert_no_flips(tot)
    else:
        for f, g in zip(frames, got):
            want = oracle.pipeline.infer_gray(states[0], None, f)
            assert np.array_equal(want, g) or (want.size == 0 and g.size == 0)


def _state_with_ids(state, rows, extra_random=0, seed=0):
    """A detector whose ids head keeps the trained rows `rows` (+ `extra_random` random rows) and the dustbin row last."""
    st = dict(state)
    w, b = state["convDb.weight"], state["convDb.bias"]
    rng = np.random.default_rng(seed)
    new_w = [w[r] for r in rows]
    new_b = [b[r] for r in rows]
    for _ in range(extra_random):
        new_w.append(rng.standard_normal(w[0].shape).astype(np.float32) * w[:16].std())
        new_b.append(np.float32(rng.standard_normal() * b[:16].std()))
    new_w.append(w[16]); new_b.append(b[16])
    st["convDb.weight"] = np.ascontiguousarray(np.stack(new_w), np.float32)
    st["convDb.bias"] = np.ascontiguousarray(np.array(new_b, np.float32))
    return st


@pytest.mark.parametrize("rows,extra", [([1, 3, 4, 6, 7, 9, 10, 12, 13, 15], 0), (list(range(16)), 8)])
def test_parity_other_n_ids(states, rows, extra):
    """n_ids = 10 (a subset of the trained id rows) and n_ids = 24 (trained rows + 8 random rows): kept set, ids and pixels against
    the oracle run on the same modified weights; dust_bin_ids = n_ids."""
    st = _state_with_ids(states[0], rows, extra, seed=3)
    n_ids = len(rows) + extra
    frames = synth.make_frames(24, 240, 320, seed=23)
    e = N.Engine(st, states[1], 240, 320, n_ids, 0, max_batch=24, max_patches=24 * 1200)
    try:
        from deepcharuco_b200.inference import _rows_to_frames
        got = _rows_to_frames(*e.infer_batch_host(frames, n_ids, True))
        got_raw = _rows_to_frames(*e.infer_batch_host(frames, n_ids, False))
    finally:
        e.close()
    reps = []
    for f, g, gr in zip(frames, got, got_raw):
        want_raw = oracle.pipeline.infer_gray(st, None, f, dust_bin_ids=n_ids)
        assert np.array_equal(want_raw, gr) or (want_raw.size == 0 and gr.size == 0)
        reps.append(parity.compare_frame((st, states[1]), f, g, gr, dust_bin_ids=n_ids))
    tot = parity.summarise(reps)
    print(f"PARITY n_ids={n_ids}:", tot)
    assert tot["K"] > 100
    parity.assert_parity(tot)
    parity.assert_no_flips(tot)


def test_predicted_handoff_is_exact(states, golden_edge, monkeypatch):
    """Sync-free detector -> RefineNet hand-off: RefineNet is enqueued for a number of 4096-patch chunks predicted from recent calls and its
    kernels take the true patch count from device memory.  Sequence sparse -> crowded (prediction far too low: the missing chunks
    are launched once the count is known) -> sparse (prediction far too high: whole chunks find no work) -> empty frames; every
    result equals the one of an engine that reads the count back first (DCU_DEVICE_COUNT=0)."""
    names = golden_edge["names"].tolist()
    crowded = golden_edge["frames"][int(np.argmax(golden_edge["counts"]))]            # 12 boards: 192 corners
    sparse = synth.make_frames(16, 240, 320, seed=31)
    batches = [sparse, np.repeat(crowded[None], 64, axis=0), sparse, np.zeros((16, 240, 320), np.uint8), sparse[:5]]
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DCU_DEVICE_COUNT", mode)
        e = N.Engine(states[0], states[1], 240, 320, 16, 0, max_batch=64, max_patches=16384)
        try:
            outs[mode] = [[a.copy() for a in e.infer_batch_host(b, 16, True)] for b in batches]
        finally:
            e.close()
    assert int(outs["1"][1][0].sum()) == 64 * int(golden_edge["counts"].max()) > 3 * 4096 - 1
    for got, want in zip(outs["1"], outs["0"]):
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    assert names
