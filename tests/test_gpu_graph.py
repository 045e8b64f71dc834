"""Small-batch CUDA-graph path of dcu_infer_batch_host (one frame per call = the reference's own benchmark loop,
src/benchmark.py:38-53): results must be identical to the kernel-by-kernel path, call after call, including frames with no
corners and frames with more corners than the graph's fixed number of patch slots."""
import os

import numpy as np
import pytest

import deepcharuco_b200 as dc
from deepcharuco_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    return dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda")


def test_graph_replay_equals_batched_path(models, golden_sample, golden_edge):
    deepc, refinenet = models
    frames = synth.make_frames(12, 240, 320, seed=11)
    want = dc.infer_batch(frames, 16, deepc, refinenet)              # n = 12 > graph_max_n: kernel by kernel
    want_raw = dc.infer_batch(frames, 16, deepc, None)
    for rep in range(3):                                             # call 1 warms, call 2 captures, call 3+ replay
        for i in range(12):
            got = dc.infer_batch(frames[i:i + 1], 16, deepc, refinenet)[0]
            assert got.shape == want[i].shape and np.array_equal(got, want[i]), (rep, i)
            raw = dc.infer_batch(frames[i:i + 1], 16, deepc, None)[0]
            assert raw.dtype == want_raw[i].dtype and np.array_equal(raw, want_raw[i])
    # n = 3 frames per call through its own graph
    for rep in range(3):
        got = dc.infer_batch(frames[3:6], 16, deepc, refinenet)
        assert all(np.array_equal(a, b) for a, b in zip(got, want[3:6]))
    kp, _ = dc.infer_image(golden_sample["bgr"], 16, deepc, refinenet)
    assert np.abs(kp - golden_sample["out_refined"]).max() <= 1e-3


def test_graph_handles_empty_and_crowded_frames(models):
    deepc, refinenet = models
    flat = np.full((1, 240, 320), 128, np.uint8)
    board = synth.make_frames(1, 240, 320, seed=4)
    # a frame tiled with boards has far more than 32 corners: the replay must fall back for the RefineNet part
    from conftest import load_golden
    edge = load_golden("edge_cases.npz")
    crowded = edge["frames"][int(np.argmax(edge["counts"]))][None]
    many = dc.infer_batch(np.concatenate([crowded] * 9), 16, deepc, refinenet)[0]      # n = 9: not a graph call
    assert many.shape[0] > 32
    for rep in range(4):
        assert dc.infer_batch(flat, 16, deepc, refinenet)[0].size == 0
        a = dc.infer_batch(board, 16, deepc, refinenet)[0]
        b = dc.infer_batch(np.concatenate([board] * 9), 16, deepc, refinenet)[0]
        assert np.array_equal(a, b)
        c = dc.infer_batch(crowded, 16, deepc, refinenet)[0]
        assert np.array_equal(c, many)
