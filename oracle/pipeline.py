"""Oracle restatement of inference.infer_image / solve_pnp (/root/reference/src/inference.py:15-70)."""
import numpy as np
import torch

from . import decode as D
from . import nets


def infer_gray(state_det, state_ref, gray_u8, dust_bin_ids=16, return_stages=False):
    """One grayscale u8 frame (H,W) through the whole path, mirroring inference.py:41-70."""
    img = D.pre_bgr_image(gray_u8)                                   # :41
    x = torch.from_numpy(img)[None]                                  # :42, net.py:97
    loc, ids = nets.detector_forward(state_det, x)                   # :43
    loc, ids = loc.numpy(), ids.numpy()
    kpts, ids_found = D.pred_to_keypoints(loc, ids, dust_bin_ids)    # :44
    stages = {"loc": loc, "ids": ids, "kpts": kpts, "ids_found": ids_found}
    if ids_found.shape[0] == 0:                                      # :51-52
        return (np.array([]), stages) if return_stages else np.array([])
    out_kpts = kpts
    if state_ref is not None:
        patches = D.extract_patches(img, kpts)                       # :55
        heat = nets.refinenet_forward(state_ref, torch.from_numpy(patches)[:, None])[:, 0].numpy()   # :57
        out_kpts, corners = D.refine_corners(heat, kpts)
        stages.update(patches=patches, heat=heat, corners=corners, refined=out_kpts)
    res = D.marshal_keypoints(out_kpts, ids_found)                   # :68-70
    return (res, stages) if return_stages else res


def infer_image(state_det, state_ref, img_bgr, dust_bin_ids=16):
    """inference.py:32-70 with draw_pred=False; BGR->gray stays cv2 (third-party fixed-point luma, :40)."""
    import cv2
    gray = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2GRAY)
    return infer_gray(state_det, state_ref, gray, dust_bin_ids)


def infer_gray_batch(state_det, state_ref, frames_u8, dust_bin_ids=16):
    """Per-frame loop, the only batching the reference supports."""
    return [infer_gray(state_det, state_ref, f, dust_bin_ids) for f in frames_u8]


def solve_pnp(keypoints, col_count, row_count, square_len, camera_matrix, dist_coeffs):
    """inference.py:15-29 -- object point of id k is ((k % (rows-1))+1, (k // (rows-1))+1, 0) * square_len."""
    import cv2
    if keypoints.shape[0] < 4:
        return False, None, None
    inn_rc = np.arange(1, row_count)
    inn_cc = np.arange(1, col_count)
    object_points = np.zeros(((col_count - 1) * (row_count - 1), 3), np.float32)
    object_points[:, :2] = np.array(np.meshgrid(inn_rc, inn_cc)).reshape((2, -1)).T * square_len
    image_points = keypoints[:, :2].astype(np.float32)
    found = object_points[keypoints[:, 2].astype(int)]
    return cv2.solvePnP(found, image_points, camera_matrix, dist_coeffs)
