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


# ---------------------------------------------------------------------------------------------------------------------
# Frames generated ON THE DEVICE (csrc/synth.cu).  The host only derives the per-frame parameters
# (a few hundred bytes per frame, vectorised over the batch); warp, paste, blur, gain and noise are one kernel launch.
# ---------------------------------------------------------------------------------------------------------------------
_STREAM_PARAMS = 1
_BLUR_R = 6
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)


def _philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on uint32 arrays; the same generator csrc/synth.cu runs per lattice node / pixel."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, np.uint32) for c in (c0, c1, c2, c3)])
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = c0.astype(np.uint64) * _PH_M0
        p1 = c2.astype(np.uint64) * _PH_M1
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint32(k0), lo1, hi0 ^ c3 ^ np.uint32(k1), lo0
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def gpu_frame_params(n, H, W, seed, n_boards=None, first_index=0, base=240):
    """Per-frame parameters of the device generator for frames first_index .. first_index + n - 1 of stream `seed`:
    a dict of arrays (lat_step, bg_lo, bg_hi, gain, blur_w (n,13), H / Hinv (n,nb,3,3): board px -> frame px and back,
    corners (n,nb,16,2): the projected inner corners = the frames' ground-truth labels, id = index 0..15)."""
    if n_boards is None:
        n_boards = 1 if (H <= 240 and W <= 320) else 4
    assert 0 <= n_boards <= 4
    idx = (first_index + np.arange(n)).astype(np.uint32)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    r = np.stack([np.stack(_philox4x32(np.uint32(j), idx, np.uint32(_STREAM_PARAMS), np.uint32(0), k0, k1), 1) for j in range(16)], 1)
    u = r.reshape(n, 64).astype(np.float64) / 4294967296.0
    p = dict(n_boards=n_boards)
    sigma_bg = 1.0 + 5.0 * u[:, 0]
    p["lat_step"] = np.maximum(4, np.rint(4.0 * sigma_bg)).astype(np.int32)
    p["bg_lo"] = np.floor(80.0 * u[:, 1]).astype(np.float32)
    p["bg_hi"] = (120.0 + np.floor(136.0 * u[:, 2])).astype(np.float32)
    sigma = 0.3 + 1.2 * u[:, 3]
    taps = np.exp(-0.5 * (np.arange(-_BLUR_R, _BLUR_R + 1)[None, :] / sigma[:, None]) ** 2)
    p["blur_w"] = (taps / taps.sum(1, keepdims=True)).astype(np.float32)
    p["gain"] = (0.3 + 0.8 * u[:, 4]).astype(np.float32)
    Hs = np.zeros((n, n_boards, 3, 3)); Hi = np.zeros((n, n_boards, 3, 3)); corners = np.zeros((n, n_boards, 16, 2))
    g = np.array(np.meshgrid(np.arange(1, 5), np.arange(1, 5))).reshape(2, -1).T * (base / 5.0)     # aruco_utils.py:122-131
    g1 = np.concatenate([g, np.ones((16, 1))], 1)
    src = np.array([[0, 0], [base - 1, 0], [base - 1, base - 1], [0, base - 1]], np.float64)
    sq0 = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64)
    for b in range(n_boards):
        v = u[:, 8 + 12 * b: 8 + 12 * (b + 1)]
        side = (0.3 + 0.6 * v[:, 0]) * base
        ang = 2.0 * np.pi * v[:, 1]
        if n_boards == 1:
            cx, cy = W / 2 + (v[:, 2] * 0.4 - 0.2) * W, H / 2 + (v[:, 3] * 0.4 - 0.2) * H
        else:
            cx, cy = (0.15 + 0.7 * v[:, 2]) * W, (0.15 + 0.7 * v[:, 3]) * H
        c, s = np.cos(ang), np.sin(ang)
        rot = np.stack([np.stack([c, -s], 1), np.stack([s, c], 1)], 1)                               # (n,2,2)
        dst = np.einsum("kj,nij->nki", sq0, rot) * (side / 2)[:, None, None] + np.stack([cx, cy], 1)[:, None, :]
        dst = dst + (v[:, 4:12].reshape(n, 4, 2) * 0.16 - 0.08) * side[:, None, None]
        A = np.zeros((n, 8, 8)); rhs = np.zeros((n, 8))
        for i in range(4):
            x, y = src[i]
            A[:, i, 0], A[:, i, 1], A[:, i, 2] = x, y, 1.0
            A[:, i, 6], A[:, i, 7] = -x * dst[:, i, 0], -y * dst[:, i, 0]
            A[:, i + 4, 3], A[:, i + 4, 4], A[:, i + 4, 5] = x, y, 1.0
            A[:, i + 4, 6], A[:, i + 4, 7] = -x * dst[:, i, 1], -y * dst[:, i, 1]
            rhs[:, i], rhs[:, i + 4] = dst[:, i, 0], dst[:, i, 1]
        h = np.linalg.solve(A, rhs[:, :, None])[:, :, 0]
        Hm = np.concatenate([h, np.ones((n, 1))], 1).reshape(n, 3, 3)
        Hs[:, b] = Hm
        Hi[:, b] = np.linalg.inv(Hm)
        q = np.einsum("kj,nij->nki", g1, Hm)
        corners[:, b] = q[:, :, :2] / q[:, :, 2:3]
    p["H"], p["Hinv"], p["corners"] = Hs, Hi, corners
    return p


def pack_frame_params(p, H, W):
    """dict of arrays -> ctypes array of DcuSynthFrame."""
    from . import _native as N
    n = len(p["lat_step"])
    arr = (N.DcuSynthFrame * n)()
    raw = np.frombuffer(arr, dtype=np.uint8).reshape(n, -1)
    rec = np.zeros(n, dtype=np.dtype([("lat", np.int32, 4), ("f", np.float32, 4), ("w", np.float32, 16), ("h", np.float64, (4, 9))]))
    assert rec.dtype.itemsize == raw.shape[1]
    rec["lat"][:, 0] = p["lat_step"]
    rec["lat"][:, 1] = H // p["lat_step"] + 3
    rec["lat"][:, 2] = W // p["lat_step"] + 3
    rec["lat"][:, 3] = p["n_boards"]
    rec["f"][:, 0], rec["f"][:, 1], rec["f"][:, 2] = p["bg_lo"], p["bg_hi"], p["gain"]
    rec["w"][:, :2 * _BLUR_R + 1] = p["blur_w"]
    rec["h"][:, :p["n_boards"]] = p["Hinv"].reshape(n, p["n_boards"], 9)
    raw[:] = rec.view(np.uint8).reshape(n, -1)
    return arr


_GPU_STATE = {}


def make_frames_gpu(n, H=240, W=320, seed=0, n_boards=None, first_index=0, device=0, return_device=False):
    """(n, H, W) uint8 frames generated on the B200 + their ground-truth corners (n, n_boards, 16, 2) float64 (ids 0..15).
    Deterministic in (seed, first_index + i, H, W, n_boards) -- frame i does not depend on n.  return_device: a CUDA tensor."""
    import torch
    from . import _native as N
    from .inference import _scratch_context
    if H % 8 or W % 8 or H < 24 or W < 24:
        raise ValueError(f"frame size {W}x{H} must be a multiple of 8 (>= 24), like the engine's")
    if n_boards is None:
        n_boards = 1 if (H <= 240 and W <= 320) else 4
    if n <= 0:
        empty = np.zeros((0, H, W), np.uint8)
        return (torch.from_numpy(empty).to(torch.device("cuda", int(device))) if return_device else empty), np.zeros((0, n_boards, 16, 2))
    p = gpu_frame_params(n, H, W, seed, n_boards, first_index)
    dev = torch.device("cuda", int(device))
    key = (int(device),)
    if key not in _GPU_STATE:
        _GPU_STATE[key] = torch.from_numpy(np.ascontiguousarray(board_render(240))).to(dev)
    board = _GPU_STATE[key]
    eng = _scratch_context(16, int(device)).engine(H, W)
    frames = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    arr = pack_frame_params(p, H, W)
    s = torch.cuda.current_stream(dev)
    N.check(N.lib().dcu_synth_frames(eng.handle, arr, n, int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_index), board.data_ptr(), 240,
                                     frames.data_ptr(), s.cuda_stream))
    s.synchronize()          # `arr` (host) is read by an asynchronous copy
    return (frames if return_device else frames.cpu().numpy()), p["corners"]


def corner_labels(corners, H, W):
    """Ground-truth label keypoints per frame in the reference's format (inference.py:150-152): rows [x, y, id] sorted by id, for
    the corners that fall inside the frame (first board only carries the ids the reference's labels use)."""
    out = []
    for c in corners:
        rows = []
        for b in range(c.shape[0]):
            for k in range(16):
                x, y = c[b, k]
                if 0 <= x < W and 0 <= y < H:
                    rows.append([x, y, float(k)])
        rows.sort(key=lambda r: r[2])
        out.append(np.array(rows, np.float64).reshape(-1, 3))
    return out


def warp_perspective_u8_gpu(src_u8, M, dsize, device=0):
    """cv2.warpPerspective(src, M, dsize, flags=cv2.INTER_LINEAR) for one uint8 single-channel image, on the B200 (bit-exact)."""
    import cv2
    import torch
    from . import _native as N
    from .inference import _scratch_context
    Wd, Hd = dsize
    src = np.ascontiguousarray(src_u8, np.uint8)
    minv = np.ascontiguousarray(cv2.invert(np.asarray(M, np.float64))[1], np.float64)       # what cv2 does internally
    dev = torch.device("cuda", int(device))
    d_src = torch.from_numpy(src).to(dev)
    dst = torch.empty((Hd, Wd), dtype=torch.uint8, device=dev)
    eng = _scratch_context(16, int(device)).engine(240, 320)
    s = torch.cuda.current_stream(dev)
    N.check(N.lib().dcu_warp_perspective_u8(eng.handle, d_src.data_ptr(), src.shape[0], src.shape[1], minv.ctypes.data, dst.data_ptr(),
                                            Hd, Wd, s.cuda_stream))
    s.synchronize()
    return dst.cpu().numpy()
