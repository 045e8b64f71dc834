"""Host-side logic of the product package that needs no GPU: BN folding, normalisation LUT,
result marshalling, synthetic-frame determinism, sharding."""
import numpy as np
import torch
import torch.nn.functional as F

import oracle
from deepcharuco_b200 import _native as N
from deepcharuco_b200 import inference as I
from deepcharuco_b200 import sharding, synth


def test_bn_fold_is_bit_exact_with_aten(states):
    """y = fma(x, alpha, beta) reproduces ATen's eval BatchNorm2d bit for bit (the CUDA epilogue applies exactly this)."""
    torch.manual_seed(0)
    for st, names in ((states[0], ["conv1a", "conv1b", "conv3a", "convPa", "convDa"]), (states[1], ["conv2a", "conv5b", "convPa"])):
        for name in names:
            alpha, beta = N.fold_bn(st, name)
            bn = "bn" + name[4:]
            c = alpha.shape[0]
            x = torch.randn(2, c, 9, 7) * 4
            want = F.batch_norm(x, torch.from_numpy(st[bn + ".running_mean"]), torch.from_numpy(st[bn + ".running_var"]),
                                torch.from_numpy(st[bn + ".weight"]), torch.from_numpy(st[bn + ".bias"]), False, 0.1, 1e-5).numpy()
            x64 = x.numpy().astype(np.float64)
            got = (x64 * alpha.astype(np.float64)[None, :, None, None] + beta.astype(np.float64)[None, :, None, None]).astype(np.float32)
            assert np.array_equal(got, want), name


def test_heads_have_no_bn(states):
    assert N.fold_bn(states[0], "convPb") == (None, None)
    arr, keep = N.layer_table(states[0], N.DET_LAYERS)
    assert [int(a.cout) for a in arr] == [64, 64, 64, 64, 128, 128, 128, 128, 256, 65, 256, 17]
    assert arr[9].alpha is None and arr[8].alpha is not None


def test_normalisation_lut_matches_reference_preprocessing():
    """The engine's 256-entry table is ((float)i - 128.0f) / 255.0f; numpy's float32 (x-128)/255 must agree exactly."""
    lut = ((np.arange(256, dtype=np.float32) - np.float32(128)) / np.float32(255)).astype(np.float32)
    img = np.arange(256, dtype=np.uint8).reshape(16, 16)
    assert np.array_equal(I.pre_bgr_image(img)[0], lut.reshape(16, 16))
    assert np.array_equal(I.pre_bgr_image(img), oracle.pre_bgr_image(img))
    # the multiply-by-reciprocal shortcut is NOT equivalent -- keep the division
    assert not np.array_equal(lut, (np.arange(256, dtype=np.float32) - 128) * np.float32(1 / 255))


def test_rows_to_frames_marshalling():
    counts = np.array([2, 0, 1], np.int32)
    offsets = np.array([0, 2, 2], np.int32)
    kpts = np.array([[10, 20, 3, 7], [11, 21, 5, 9], [1, 2, 0, 4]], np.int32)
    refined = np.array([[10.125, 19.875], [11.5, 21.0], [0.875, 2.25]], np.float32)
    out = I._rows_to_frames(counts, offsets, kpts, refined)
    assert out[0].dtype == np.float64 and out[0].tolist() == [[10.125, 19.875, 3.0], [11.5, 21.0, 5.0]]
    assert out[1].shape == (0,) and out[2].tolist() == [[0.875, 2.25, 0.0]]
    raw = I._rows_to_frames(counts, offsets, kpts, None)
    assert raw[0].dtype == np.int64 and raw[0].tolist() == [[10, 20, 3], [11, 21, 5]]


def test_marshal_matches_reference_sort():
    kp = np.array([[5, 5], [1, 1], [9, 9], [3, 3]])
    ids = np.array([7, 2, 7, 0])
    out = oracle.marshal_keypoints(kp, ids)
    assert out.tolist() == [[3, 3, 0], [1, 1, 2], [5, 5, 7], [9, 9, 7]]      # stable: duplicates keep row-major order


def test_synth_is_deterministic_and_has_boards(golden_synth):
    frames = synth.make_frames(16, 240, 320, seed=0)
    assert frames.dtype == np.uint8 and np.array_equal(frames, golden_synth["frames"])
    assert not np.array_equal(synth.make_frames(2, seed=1), frames[:2])
    assert synth.tile_frames(frames, 40).shape == (40, 240, 320)


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 2048):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_round_trip():
    res = [np.array([[1.5, 2.5, 3.0]]), np.array([]), np.array([[0.0, 1.0, 2.0], [4.0, 5.0, 6.0]])]
    c, f = sharding.pack_results(res)
    back = sharding.unpack_results(c, f)
    assert all(np.array_equal(a, b) for a, b in zip(res, back))


def test_bgr_to_gray_fixed_point_matches_cv2():
    """The device kernel's formula (csrc/decode.cu: bgr_to_gray_kernel) is OpenCV's 8-bit luma; pin it against cv2 itself."""
    import cv2
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (240, 320, 3)).astype(np.uint8)
    img[:16, :16] = np.array([[b, g, 255 - b] for b in range(16) for g in range(0, 256, 16)], np.uint8).reshape(16, 16, 3)
    B, G, R = (img[..., i].astype(np.int64) for i in range(3))
    got = ((3735 * B + 19235 * G + 9798 * R + 16384) >> 15).astype(np.uint8)
    assert np.array_equal(got, cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
