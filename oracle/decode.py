"""Oracle restatement of the integer decode / gather / argmax steps (numpy).

Each function cites the reference lines it follows
(/root/reference/src/models/model_utils.py, refinenet.py, inference.py).
"""
import numpy as np


def pre_bgr_image(gray_u8):
    """model_utils.py:46-50 -- float32(x), (x-128)/255 (true fp32 division), add channel dim."""
    image = gray_u8[..., np.newaxis].astype(np.float32)
    image = (image - 128) / 255
    return image.transpose((2, 0, 1))


def pred_argmax(loc_hat, ids_hat, dust_bin_ids):
    """model_utils.py:72-78 -- first-max argmax over channels; ids := dustbin where loc argmax == 64."""
    assert loc_hat.ndim == 4 and ids_hat.ndim == 4
    ids_argmax = np.argmax(ids_hat, axis=1)          # np.argmax, like torch.argmax, returns the first max
    loc_argmax = np.argmax(loc_hat, axis=1)
    ids_argmax = np.where(loc_argmax == 64, dust_bin_ids, ids_argmax)   # 64 hard-coded, model_utils.py:77
    return loc_argmax.astype(np.int64), ids_argmax.astype(np.int64)


def label_to_keypoints(loc, ids, dust_bin_ids):
    """model_utils.py:108-124 -- row-major nonzero of (ids != dustbin); x = 8*col + p%8, y = 8*row + p//8.

    The batch index is discarded exactly as in the reference (:121-122 use
    indices[:, -1] / indices[:, -2] only)."""
    assert loc.ndim == 3 and ids.ndim == 3
    mask = ids != dust_bin_ids
    indices = np.argwhere(mask)                      # row-major, like torch.nonzero
    ids_found = ids[mask]
    region_pixel = loc[mask]
    xs = 8 * indices[:, -1] + (region_pixel % 8)
    ys = 8 * indices[:, -2] + (region_pixel // 8)
    return np.stack((xs, ys), axis=1).astype(np.int64).reshape(-1, 2), ids_found.astype(np.int64)


def pred_to_keypoints(loc_hat, ids_hat, dust_bin_ids):
    """model_utils.py:81-88."""
    loc_argmax, ids_argmax = pred_argmax(loc_hat, ids_hat, dust_bin_ids)
    return label_to_keypoints(loc_argmax, ids_argmax, dust_bin_ids)


def extract_patches(img, keypoints, patch_size=24):
    """model_utils.py:19-36 -- zero-pad the NORMALISED image by 12, take rows y..y+23 / cols x..x+23 of the
    padded image, i.e. the window [y-12, y+12) x [x-12, x+12) of the original.  img: (1,H,W) float32."""
    pad = patch_size // 2
    padded = np.pad(img[0], ((pad, pad), (pad, pad)), mode="constant", constant_values=0)
    k = keypoints.shape[0]
    out = np.empty((k, patch_size, patch_size), np.float32)
    for i in range(k):
        x, y = int(keypoints[i, 0]), int(keypoints[i, 1])
        out[i] = padded[y:y + patch_size, x:x + patch_size]
    return out


def bargmax2d(heat):
    """model_utils.py:39-43 (speedy_bargmax2d) -- first max of the flattened map -> (col, row)."""
    k, h, w = heat.shape
    idx = np.argmax(heat.reshape(k, -1), axis=1)
    return np.stack((idx % w, idx // w), axis=1).astype(np.int64)


def refine_corners(heat, keypoints):
    """refinenet.py:108-114 -- corners_og = (corners - 32) / 8 + keypoints, evaluated in float32
    (int64 true-divide -> float32, + int64 -> float32, as torch type promotion does)."""
    corners = bargmax2d(heat)
    corners_og = ((corners - 32).astype(np.float32) / np.float32(8)) + keypoints.astype(np.float32)
    return corners_og.astype(np.float32), corners


def marshal_keypoints(keypoints, ids_found):
    """inference.py:68-70 -- stable sort by id, rows [x, y, id]; float64 when refined, int64 when raw."""
    order = sorted(range(len(ids_found)), key=lambda i: ids_found[i])   # Python sorted is stable, as in the reference
    return np.array([[keypoints[i][0], keypoints[i][1], ids_found[i]] for i in order])


def resize_linear_u8(src, dsize):
    """cv2.resize(src, (W, H), interpolation=cv2.INTER_LINEAR) for uint8 (H,W) or (H,W,C) when shrinking -- the call of the reference's
    evaluation loop (inference.py:131-132).  Third-party arithmetic (OpenCV imgproc/resize.cpp, 8-bit fixed-point path): per axis the
    source index and two 11-bit weights from a float32 fraction of (d + 0.5) * scale - 0.5 (clamped at the borders), horizontal pass
    in int, vertical pass (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.  Pinned against cv2 in
    tests/test_oracle_synth.py."""
    Wd, Hd = dsize
    s = src if src.ndim == 3 else src[..., None]
    Hs, Ws, _ = s.shape

    def coeffs(n_dst, n_src):
        scale = np.float64(n_src) / n_dst
        d = np.arange(n_dst)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        i = np.floor(f).astype(np.int64)
        f = (f - i.astype(np.float32)).astype(np.float32)
        lo = i < 0
        f = np.where(lo, np.float32(0), f); i = np.where(lo, 0, i)
        hi = i >= n_src - 1
        f = np.where(hi, np.float32(0), f); i = np.where(hi, n_src - 1, i)
        a0 = np.clip(np.rint((np.float32(1.0) - f) * np.float32(2048)), -32768, 32767).astype(np.int64)
        a1 = np.clip(np.rint(f * np.float32(2048)), -32768, 32767).astype(np.int64)
        return i, a0, a1

    sx, ax0, ax1 = coeffs(Wd, Ws)
    sy, ay0, ay1 = coeffs(Hd, Hs)
    sx1, sy1 = np.minimum(sx + 1, Ws - 1), np.minimum(sy + 1, Hs - 1)
    S = s.astype(np.int64)
    r0 = S[sy][:, sx] * ax0[None, :, None] + S[sy][:, sx1] * ax1[None, :, None]
    r1 = S[sy1][:, sx] * ax0[None, :, None] + S[sy1][:, sx1] * ax1[None, :, None]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[..., 0]
