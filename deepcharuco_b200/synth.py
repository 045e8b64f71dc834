"""Seeded synthetic ChArUco frames for parity tests and benchmarks.

Frames must contain boards -- uniform noise decodes to K=0 corners and the
RefineNet leg would do no work.  The recipe follows the augmentation ranges the
reference trains with (/root/reference/src/transformations.py:22-52,105-114:
board side 0.25-0.9 of min(H,W), any rotation, blur, brightness) and the board
of src/demo_config.yaml:2-6 (5x5, DICT_4X4_50, square 0.01, marker 0.0075).
Host-side numpy/cv2 only; runs identically on the build box and the GPU box.
"""
import numpy as np

_BOARD_CACHE = {}


def board_render(side=240):
    """The 5x5 DICT_4X4_50 board rendered at side x side (aruco_utils.py:53-73,122-125)."""
    import cv2
    if side not in _BOARD_CACHE:
        d = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50)
        b = cv2.aruco.CharucoBoard((5, 5), 0.01, 0.0075, d)
        _BOARD_CACHE[side] = b.generateImage((side, side))
    return _BOARD_CACHE[side]


def _one_frame(rng, H, W, n_boards=1, base=240):
    import cv2
    # background: blurred uniform noise, min-max normalised to a random [lo, hi]
    bg = rng.integers(0, 256, (H, W)).astype(np.float32)
    bg = cv2.GaussianBlur(bg, (0, 0), float(rng.uniform(1.0, 6.0)))
    lo, hi = float(rng.integers(0, 80)), float(rng.integers(120, 256))
    bg = (bg - bg.min()) / max(float(bg.max() - bg.min()), 1e-6) * (hi - lo) + lo
    frame = bg
    board = board_render(240).astype(np.float32)
    for b in range(n_boards):
        side = float(rng.uniform(0.3, 0.9)) * base
        ang = float(rng.uniform(0, 2 * np.pi))
        if n_boards == 1:
            cx = W / 2 + float(rng.uniform(-0.2, 0.2)) * W
            cy = H / 2 + float(rng.uniform(-0.2, 0.2)) * H
        else:
            cx = float(rng.uniform(0.15, 0.85)) * W
            cy = float(rng.uniform(0.15, 0.85)) * H
        c, s = np.cos(ang), np.sin(ang)
        sq = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float32) * (side / 2)
        dst = sq @ np.array([[c, -s], [s, c]], np.float32).T + np.array([cx, cy], np.float32)
        dst += rng.uniform(-0.08, 0.08, (4, 2)).astype(np.float32) * side
        src = np.array([[0, 0], [239, 0], [239, 239], [0, 239]], np.float32)
        M = cv2.getPerspectiveTransform(src, dst.astype(np.float32))
        warped = cv2.warpPerspective(board, M, (W, H), flags=cv2.INTER_LINEAR)
        mask = cv2.warpPerspective(np.ones_like(board), M, (W, H), flags=cv2.INTER_LINEAR)
        frame = frame * (1 - mask) + warped * mask
    frame = cv2.GaussianBlur(frame, (0, 0), float(rng.uniform(0.3, 1.5)))
    frame = frame * float(rng.uniform(0.3, 1.1)) + rng.normal(0, 3, (H, W)).astype(np.float32)
    return np.clip(np.rint(frame), 0, 255).astype(np.uint8)


def make_frames(n, H=240, W=320, seed=0, n_boards=None):
    """(n, H, W) uint8 grayscale frames, each with >=1 warped board.  Deterministic in (n, H, W, seed).

    640x480 frames keep the board at the trained scale (side <= 216 px) and carry
    4 boards so K stays near 4x the 320x240 figure."""
    rng = np.random.default_rng(seed)
    if n_boards is None:
        n_boards = 1 if (H <= 240 and W <= 320) else 4
    out = np.empty((n, H, W), np.uint8)
    for i in range(n):
        out[i] = _one_frame(rng, H, W, n_boards)
    return out


def tile_frames(frames, n):
    """Cycle a small pool of distinct frames up to n (benchmarks; generation cost stays bounded)."""
    reps = (n + len(frames) - 1) // len(frames)
    return np.ascontiguousarray(np.concatenate([frames] * reps, 0)[:n])
