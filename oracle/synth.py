"""Oracle for the device frame generator -- TEST INFRASTRUCTURE ONLY.

Two parts:
  * `warp_perspective_u8`: a restatement of cv2.warpPerspective(src u8, M, dsize, INTER_LINEAR, BORDER_CONSTANT 0) -- the third-party
    routine (OpenCV imgproc, imgwarp.cpp: warpPerspectiveInvoker + remapBilinear, fixed point with INTER_BITS = 5 and 15-bit
    weights) that the reference's board augmentation ends in (albumentations A.Affine -> cv2.warpAffine / warpPerspective,
    /root/reference/src/transformations.py:33-35) and that deepcharuco_b200/synth.py calls.  Pinned against cv2 itself in
    tests/test_oracle_synth.py (bit-exact on random homographies).
  * `make_frames`: the numpy mirror of csrc/synth.cu -- same counter-based random numbers (Philox4x32-10), same float32 operations
    in the same order (no fused multiply-adds, no transcendental functions on the device: blur weights and homographies are
    per-frame parameters computed here in float64), so the CUDA frames can be compared bit for bit.
The recipe (board side 0.3-0.9 x 240 px, any rotation, corner jitter, blur, gain, noise) is that of deepcharuco_b200/synth.py,
i.e. the augmentation ranges of transformations.py:22-52,105-114; the background is interpolated lattice noise instead of blurred
white noise (no global min/max pass).
"""
import numpy as np

# ---------------------------------------------------------------------------------------------------------------------
# cv2.warpPerspective, u8, INTER_LINEAR, BORDER_CONSTANT(0)
# ---------------------------------------------------------------------------------------------------------------------
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS


def bilinear_tab_i():
    """OpenCV's BilinearTab_i (imgwarp.cpp initInterTab2D, fixpt): weights (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx of the 1/32
    steps, float32 products scaled by 2^15 and saturated to int16.  Only alpha = 0 saturates (32768 -> 32767); OpenCV's fix-up of
    the missing 1 walks entries 3..6 of the 2x2 block (its loop is written for larger kernels), so entry 3 receives it."""
    t = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, 4), np.int32)
    one = np.float32(1.0)
    for i in range(INTER_TAB_SIZE):
        fy = np.float32(i) * np.float32(1.0 / INTER_TAB_SIZE)
        ty = (one - fy, fy)
        for j in range(INTER_TAB_SIZE):
            fx = np.float32(j) * np.float32(1.0 / INTER_TAB_SIZE)
            tx = (one - fx, fx)
            isum = 0
            for k1 in range(2):
                for k2 in range(2):
                    v = np.float32(ty[k1] * tx[k2]) * np.float32(COEF_SCALE)
                    iv = int(np.clip(np.rint(v), -32768, 32767))
                    t[i, j, k1 * 2 + k2] = iv
                    isum += iv
            if isum != COEF_SCALE:
                t[i, j, 3] -= isum - COEF_SCALE
    return t


_TAB = None


def warp_coords(minv, W, H):
    """Fixed-point source coordinates of every destination pixel, as warpPerspectiveInvoker computes them: float64, blocks of
    64 columns (X0 is evaluated at the block's first column, then + M0 * x1), 1/32-pixel units, round half to even."""
    m = np.asarray(minv, np.float64).reshape(-1)
    bh0 = min(16, H); bw0 = min(1024 // bh0, W)
    ys = np.arange(H, dtype=np.float64)[:, None]
    xs = np.arange(W)[None, :]
    xb = (xs // bw0) * bw0
    x1 = (xs - xb).astype(np.float64)
    xb = xb.astype(np.float64)
    X0 = (m[0] * xb + m[1] * ys) + m[2]
    Y0 = (m[3] * xb + m[4] * ys) + m[5]
    W0 = (m[6] * xb + m[7] * ys) + m[8]
    Wv = W0 + m[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        Wv = np.where(Wv != 0, INTER_TAB_SIZE / Wv, 0.0)
    fX = np.maximum(-2147483648.0, np.minimum(2147483647.0, (X0 + m[0] * x1) * Wv))
    fY = np.maximum(-2147483648.0, np.minimum(2147483647.0, (Y0 + m[3] * x1) * Wv))
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    return X, Y


def sample_u8(src, X, Y, constant_src=None):
    """remapBilinear (u8, 1 channel, BORDER_CONSTANT 0) at fixed-point coordinates X, Y.  constant_src: sample a virtual image that is
    `constant_src` everywhere inside src's extent (the paste mask) instead of src."""
    global _TAB
    if _TAB is None:
        _TAB = bilinear_tab_i()
    sh, sw = src.shape
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    w = _TAB[Y & (INTER_TAB_SIZE - 1), X & (INTER_TAB_SIZE - 1)]

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
        if constant_src is not None:
            return np.where(ok, np.int64(constant_src), 0)
        return np.where(ok, src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64), 0)

    acc = tap(sy, sx) * w[..., 0] + tap(sy, sx + 1) * w[..., 1] + tap(sy + 1, sx) * w[..., 2] + tap(sy + 1, sx + 1) * w[..., 3]
    return np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)


def warp_perspective_u8(src, minv, dsize, constant_src=None):
    """dst(x, y) = src(minv * (x, y, 1)) like cv2.warpPerspective(src, M, dsize, flags=INTER_LINEAR) with minv = cv2.invert(M)
    (cv2 inverts M itself; pass flags | WARP_INVERSE_MAP to cv2 to hand it minv directly)."""
    W, H = dsize
    X, Y = warp_coords(minv, W, H)
    return sample_u8(np.asarray(src, np.uint8), X, Y, constant_src)


# ---------------------------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), vectorised over counters
# ---------------------------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """All arguments uint32 arrays (broadcastable) -> four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


# ---------------------------------------------------------------------------------------------------------------------
# frame recipe
# ---------------------------------------------------------------------------------------------------------------------
STREAM_PARAMS, STREAM_LATTICE, STREAM_NOISE = 1, 2, 3
BLUR_R = 6                      # taps -6..6: cv2.GaussianBlur's kernel for float images at sigma = 1.5 (8 * sigma + 1)
BOARD_PX = 240
LATTICE_MAX = 96                # lattice values per row (W / S + 3 <= 96 for W <= 1280 at S >= 14 ...) -- see lattice_shape


def _u01(u):                    # uint32 -> float64 in [0, 1)
    return np.asarray(u, np.float64) / 4294967296.0


def frame_params(seed, index, H, W, n_boards, base=BOARD_PX):
    """Per-frame parameter block (float64 / int), derived from Philox(counter = (j, index, STREAM_PARAMS, 0), key = seed):
    background lattice step + range, per board the homography board px -> frame px, blur weights, gain."""
    seed = int(seed)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    r = np.concatenate([np.stack(philox4x32(np.uint32(j), np.uint32(index), np.uint32(STREAM_PARAMS), np.uint32(0), k0, k1)) for j in range(16)])
    u = _u01(r)                                                       # 64 uniforms
    p = {}
    sigma_bg = 1.0 + 5.0 * u[0]
    p["lat_step"] = int(max(4, round(4.0 * sigma_bg)))                # lattice spacing in pixels
    p["bg_lo"] = np.float32(np.floor(80.0 * u[1]))
    p["bg_hi"] = np.float32(120.0 + np.floor(136.0 * u[2]))
    sigma = 0.3 + 1.2 * u[3]
    taps = np.exp(-0.5 * (np.arange(-BLUR_R, BLUR_R + 1) / sigma) ** 2)
    p["blur_w"] = (taps / taps.sum()).astype(np.float32)
    p["gain"] = np.float32(0.3 + 0.8 * u[4])
    p["H"], p["Hinv"], p["corners"] = [], [], []
    for b in range(n_boards):
        v = u[8 + 12 * b: 8 + 12 * (b + 1)]
        side = (0.3 + 0.6 * v[0]) * base
        ang = 2.0 * np.pi * v[1]
        if n_boards == 1:
            cx, cy = W / 2 + (v[2] * 0.4 - 0.2) * W, H / 2 + (v[3] * 0.4 - 0.2) * H
        else:
            cx, cy = (0.15 + 0.7 * v[2]) * W, (0.15 + 0.7 * v[3]) * H
        c, s = np.cos(ang), np.sin(ang)
        sq = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64) * (side / 2)
        dst = sq @ np.array([[c, -s], [s, c]]).T + np.array([cx, cy])
        dst += (v[4:12].reshape(4, 2) * 0.16 - 0.08) * side
        src = np.array([[0, 0], [base - 1, 0], [base - 1, base - 1], [0, base - 1]], np.float64)
        Hm = homography(src, dst)
        p["H"].append(Hm)
        p["Hinv"].append(np.linalg.inv(Hm))
        # inner corners of the 5x5 board in board pixels (aruco_utils.py:122-131: meshgrid(1..rows-1, 1..cols-1) * base / 5), id = k
        g = np.array(np.meshgrid(np.arange(1, 5), np.arange(1, 5))).reshape(2, -1).T * (base / 5.0)
        q = np.concatenate([g, np.ones((16, 1))], 1) @ Hm.T
        p["corners"].append(q[:, :2] / q[:, 2:3])
    return p


def homography(src, dst):
    """3x3 H (h33 = 1) with dst ~ H src for four point pairs: the 8x8 system of cv2.getPerspectiveTransform, solved in float64."""
    A, b = np.zeros((8, 8)), np.zeros(8)
    for i in range(4):
        x, y = src[i]; u, v = dst[i]
        A[i] = [x, y, 1, 0, 0, 0, -x * u, -y * u]; b[i] = u
        A[i + 4] = [0, 0, 0, x, y, 1, -x * v, -y * v]; b[i + 4] = v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


def lattice_shape(H, W, step):
    return H // step + 3, W // step + 3


def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def make_frame(board_u8, seed, index, H, W, n_boards, params=None):
    """One frame, float32 arithmetic in the kernel's order.  Returns (frame u8 (H,W), corners float64 (n_boards,16,2)).
    params: use these per-frame parameters (same keys as frame_params) instead of deriving them -- the per-pixel pipeline is what
    is compared bit for bit; the parameters are float64 host arithmetic on either side and agree to ~1e-13."""
    f32 = np.float32
    p = frame_params(seed, index, H, W, n_boards) if params is None else params
    seed = int(seed)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    S = p["lat_step"]
    lh, lw = lattice_shape(H, W, S)
    li = np.arange(lh * lw, dtype=np.uint32)
    lat = (philox4x32(li, np.uint32(index), np.uint32(STREAM_LATTICE), np.uint32(0), k0, k1)[0] >> np.uint32(24)).astype(f32).reshape(lh, lw)
    R = BLUR_R
    ys = _reflect101(np.arange(-R, H + R), H)           # composite evaluated at reflected coordinates (BORDER_REFLECT_101)
    xs = _reflect101(np.arange(-R, W + R), W)
    # background: bilinear interpolation of the lattice, then affine map to [lo, hi]
    gy, gx = ys // S, xs // S
    fy = ((ys - gy * S).astype(f32) / f32(S))[:, None]
    fx = ((xs - gx * S).astype(f32) / f32(S))[None, :]
    a = lat[gy][:, gx]; b = lat[gy][:, gx + 1]; c = lat[gy + 1][:, gx]; d = lat[gy + 1][:, gx + 1]
    top = a + (b - a) * fx
    bot = c + (d - c) * fx
    t = top + (bot - top) * fy
    comp = p["bg_lo"] + (p["bg_hi"] - p["bg_lo"]) * (t * f32(1.0 / 255.0))
    comp = comp.astype(f32)
    # boards, pasted in order: exact cv2 fixed-point warp of texture and mask at the (unreflected -> reflected) pixel coordinates
    for bidx in range(n_boards):
        X, Y = warp_coords(p["Hinv"][bidx], W, H)
        Xr, Yr = X[ys][:, xs], Y[ys][:, xs]
        wv = sample_u8(board_u8, Xr, Yr).astype(f32)
        mk = sample_u8(board_u8, Xr, Yr, constant_src=255).astype(f32) * f32(1.0 / 255.0)
        comp = (comp * (f32(1.0) - mk) + wv * mk).astype(f32)
    # separable blur: horizontal then vertical, taps accumulated left to right / top to bottom
    w = p["blur_w"]
    hb = np.zeros((H + 2 * R, W), f32)
    for k in range(2 * R + 1):
        hb = (hb + w[k] * comp[:, k:k + W]).astype(f32)
    vb = np.zeros((H, W), f32)
    for k in range(2 * R + 1):
        vb = (vb + w[k] * hb[k:k + H, :]).astype(f32)
    # gain + noise: Irwin-Hall sum of four bytes, scaled to sigma = 3
    pi = np.arange(H * W, dtype=np.uint32)
    r0 = philox4x32(pi, np.uint32(index), np.uint32(STREAM_NOISE), np.uint32(0), k0, k1)[0]
    s4 = ((r0 & np.uint32(255)) + ((r0 >> np.uint32(8)) & np.uint32(255)) + ((r0 >> np.uint32(16)) & np.uint32(255)) + (r0 >> np.uint32(24))).astype(np.int32)
    noise = ((s4 - 510).astype(f32) * f32(3.0 / 147.8005413)).reshape(H, W)
    out = (vb * p["gain"] + noise).astype(f32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8), np.stack(p["corners"])


def make_frames(board_u8, n, H, W, seed, n_boards=None, first_index=0):
    if n_boards is None:
        n_boards = 1 if (H <= 240 and W <= 320) else 4
    frames = np.empty((n, H, W), np.uint8)
    corners = np.empty((n, n_boards, 16, 2), np.float64)
    for i in range(n):
        frames[i], corners[i] = make_frame(board_u8, seed, first_index + i, H, W, n_boards)
    return frames, corners
