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
