"""Drop-in replacement for the reference's Python inference surface.

Same names, positional order, defaults and return types as
/root/reference/src/inference.py: `load_models` (:73), `infer_image` (:32), `solve_pnp` (:15); the helper
names the reference imports next to them (`pred_to_keypoints`, `extract_patches`, `pre_bgr_image`,
model_utils.py) and duck-typed model handles (`deepc.infer_image`, `refinenet.infer_patches`,
net.py:127, refinenet.py:143).  New and additive: `infer_batch` (the reference is one frame per call).

Everything numerical runs in the hand-written sm_100a kernels behind the C ABI
(include/deepcharuco_b200.h); PyTorch is used only for device memory handed to those kernels.
There is no CPU / PyTorch fallback: without the built library or a B200 the calls raise.
"""
import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _native as N
from . import weights_io

__all__ = ["load_models", "infer_image", "infer_batch", "solve_pnp", "pred_to_keypoints", "extract_patches",
           "pre_bgr_image", "draw_inner_corners", "resize_gpu"]


def _default_conv_impl():
    v = os.environ.get("DCU_CONV_IMPL", "").lower()
    if v in ("ffma", "0"):
        return N.CONV_FFMA
    if v in ("tcgen05", "tc", "1"):
        return N.CONV_TCGEN05
    return N.CONV_DEFAULT


def _device_index(device):
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    s = str(device)
    if s in ("cuda", "cpu", "mps"):       # the reference passes 'cpu'/'cuda'/'mps'; this engine is always the B200
        return int(os.environ.get("LOCAL_RANK", "0")) if s == "cuda" and "LOCAL_RANK" in os.environ else 0
    if s.startswith("cuda:"):
        return int(s.split(":")[1])
    return 0


class _Context:
    """Weights + a cache of C engines keyed by frame size (an engine's workspace is shape-specific)."""

    def __init__(self, state_det, state_ref, n_ids, device):
        self.state_det, self.state_ref = state_det, state_ref
        self.n_ids, self.device = int(n_ids), int(device)
        self.conv_impl = _default_conv_impl()
        self._engines = {}
        _CONTEXTS.append(self)

    def engine(self, height, width, max_batch=1, max_patches=None) -> N.Engine:
        key = (int(height), int(width))
        eng = self._engines.get(key)
        want_p = max_patches if max_patches is not None else max(256, 64 * max_batch)
        if eng is None or eng.max_batch < max_batch or eng.max_patches < want_p:
            if eng is not None:
                max_batch = max(max_batch, eng.max_batch)
                want_p = max(want_p, eng.max_patches)
                eng.close()
            eng = N.Engine(self.state_det, self.state_ref, height, width, self.n_ids, self.device,
                           max_batch=max_batch, max_patches=want_p, conv_impl=self.conv_impl)
            self._engines[key] = eng
        return eng

    def set_conv_impl(self, impl):
        self.conv_impl = int(impl)
        for e in self._engines.values():
            e.set_conv_impl(impl)

    def close(self):
        for e in self._engines.values():
            e.close()
        self._engines.clear()


def _torch():
    import torch
    return torch


class DeepcHandle:
    """Stands in for `lModel` (net.py:118-128).  `infer_image(img_gray)` keeps the reference signature."""

    def __init__(self, ctx: _Context):
        self._ctx = ctx
        self.n_ids = ctx.n_ids

    def eval(self):
        return self

    def to(self, device):
        return self

    def infer_image(self, img):
        """img: (1,H,W) float32 normalised image (tensor or ndarray) -> (loc (1,65,H/8,W/8), ids (1,n_ids+1,H/8,W/8))
        CUDA tensors.  net.py:82-99 / :127-128."""
        torch = _torch()
        dev = torch.device("cuda", self._ctx.device)
        x = torch.as_tensor(img, dtype=torch.float32).to(dev).contiguous()
        assert x.ndim == 3 and x.shape[0] == 1, "expected a (1,H,W) image as the reference does"
        H, W = int(x.shape[1]), int(x.shape[2])
        eng = self._ctx.engine(H, W)
        loc = torch.empty((1, 65, H // 8, W // 8), dtype=torch.float32, device=dev)
        ids = torch.empty((1, self.n_ids + 1, H // 8, W // 8), dtype=torch.float32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().dcu_detector_forward_f32(eng.handle, x.data_ptr(), 1, loc.data_ptr(), ids.data_ptr(), s))
        return loc, ids

    __call__ = infer_image


class RefineHandle:
    """Stands in for `lRefineNet` (refinenet.py:134-145)."""

    def __init__(self, ctx: _Context):
        self._ctx = ctx

    def eval(self):
        return self

    def to(self, device):
        return self

    def infer_patches(self, patches, keypoints):
        """patches (K,24,24) or (K,1,24,24) float32, keypoints (K,2) int -> (corners_og (K,2) float32,
        corners (K,2) int64), CUDA tensors.  refinenet.py:85-115."""
        torch = _torch()
        dev = torch.device("cuda", self._ctx.device)
        p = torch.as_tensor(patches, dtype=torch.float32).to(dev)
        assert tuple(p.shape[-2:]) == (24, 24)                          # refinenet.py:102
        if p.ndim == 4:
            p = p[:, 0]
        p = p.contiguous()
        k = int(p.shape[0])
        xy = torch.as_tensor(keypoints).to(dev).to(torch.int32).contiguous()
        eng = self._ctx.engine(self._any_size()[0], self._any_size()[1], max_patches=max(256, k))
        corners = torch.empty((k, 2), dtype=torch.int32, device=dev)
        refined = torch.empty((k, 2), dtype=torch.float32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().dcu_refine_forward(eng.handle, p.data_ptr(), xy.data_ptr(), 2, k, corners.data_ptr(),
                                           refined.data_ptr(), None, s))
        return refined, corners.to(torch.int64)

    def _any_size(self):
        # RefineNet work is frame-size independent; reuse any engine, else make a small one
        for key in self._ctx._engines:
            return key
        return (240, 320)


def load_models(deepc_ckpt: str, refinenet_ckpt: Optional[str] = None, n_ids: int = 16, device='cuda'):
    """inference.py:73-84.  Accepts the reference's Lightning .ckpt files or the converted .npz."""
    state_det = weights_io.load_state(deepc_ckpt)
    state_ref = weights_io.load_state(refinenet_ckpt) if refinenet_ckpt is not None else None
    N.lib()   # fail now, loudly, if the CUDA library is not built
    ctx = _Context(state_det, state_ref, n_ids, _device_index(device))
    return DeepcHandle(ctx), (RefineHandle(ctx) if state_ref is not None else None)


def pre_bgr_image(image):
    """model_utils.py:46-50 -- host numpy in the reference too: float32, (x-128)/255, channel first."""
    image = image[..., np.newaxis].astype(np.float32)
    image = (image - 128) / 255
    return image.transpose((2, 0, 1))


def _rows_to_frames(counts, offsets, kpts, refined):
    """Packed engine output -> per-frame arrays in the reference's format (inference.py:68-70):
    float64 (K,3) [x, y, id] with RefineNet, int64 (K,3) without, np.array([]) when K == 0 (:51-52).
    One conversion for the whole batch; the per-frame arrays are slices of that fresh array (rows of a frame are contiguous)."""
    counts = np.asarray(counts)
    offsets = np.asarray(offsets)
    total = int(counts.sum())
    if refined is not None:
        rows = np.empty((total, 3), np.float64)
        rows[:, :2] = refined[:total]             # float32 -> float64, exact
        rows[:, 2] = kpts[:total, 2]
    else:
        rows = kpts[:total, :3].astype(np.int64)
    empty = np.array([])
    return [rows[o:o + c] if c else empty.copy() for c, o in zip(counts.tolist(), offsets.tolist())]


def _infer_batch_resized(frames, input_size, dust_bin_ids, deepc, refinenet):
    """Camera frames larger than the network input: cv2.resize(frame, input_size, cv2.INTER_LINEAR) (inference.py:131-132) and
    cv2.cvtColor(BGR2GRAY) (:40) both on the device (`dcu_resize_u8`, `dcu_bgr_to_gray`: bit-exact with cv2), then the pipeline."""
    torch = _torch()
    W, H = int(input_size[0]), int(input_size[1])
    n, Hs, Ws = frames.shape[:3]
    ch = 3 if frames.ndim == 4 else 1
    ctx = deepc._ctx
    eng = ctx.engine(H, W, max_batch=n)
    dev = torch.device("cuda", ctx.device)
    s = torch.cuda.current_stream(dev).cuda_stream
    src = torch.from_numpy(np.ascontiguousarray(frames)).to(dev)
    small = torch.empty((n, H, W, ch) if ch == 3 else (n, H, W), dtype=torch.uint8, device=dev)
    N.check(N.lib().dcu_resize_u8(eng.handle, src.data_ptr(), n, Hs, Ws, ch, small.data_ptr(), s))
    gray = small
    if ch == 3:
        gray = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
        N.check(N.lib().dcu_bgr_to_gray(eng.handle, small.data_ptr(), n, gray.data_ptr(), s))
    while True:
        try:
            o = eng.infer_batch_device(gray.data_ptr(), n, dust_bin_ids, refinenet is not None, s)
        except N.CapacityError:
            eng = ctx.engine(H, W, max_batch=n, max_patches=2 * eng.max_patches)
            continue
        total = int(o["total"].item())
        if total <= eng.max_patches:
            break
        eng = ctx.engine(H, W, max_batch=n, max_patches=max(total, 2 * eng.max_patches))
    counts, offsets = o["counts"][:n].cpu().numpy(), o["offsets"][:n].cpu().numpy()
    kpts = o["kpts"][:total].cpu().numpy()
    refined = o["refined"][:total].cpu().numpy() if refinenet is not None else None
    return _rows_to_frames(counts, offsets, kpts, refined)


def infer_batch(frames, dust_bin_ids: int, deepc: DeepcHandle, refinenet: Optional[RefineHandle] = None, input_size=None):
    """Batched form of infer_image: frames (N,H,W) uint8 grayscale or (N,H,W,3) BGR -> list of N keypoint arrays.

    One H2D copy of the u8 frames, the fused GPU pipeline, one D2H copy of the packed result.  input_size = (W, H): frames of
    another (larger) size are first resized to the network input on the device like cv2.resize(frame, (W, H), cv2.INTER_LINEAR)
    (the reference's evaluation loop, inference.py:131-132), bit-exact with cv2 when shrinking."""
    frames = np.asarray(frames)
    assert frames.dtype == np.uint8 and (frames.ndim == 3 or (frames.ndim == 4 and frames.shape[3] == 3)), \
        "frames must be (N,H,W) grayscale or (N,H,W,3) BGR uint8"
    if input_size is not None and (frames.shape[2], frames.shape[1]) != (int(input_size[0]), int(input_size[1])):
        if frames.shape[0] == 0:
            return []
        return _infer_batch_resized(frames, input_size, dust_bin_ids, deepc, refinenet)
    n, H, W = frames.shape[:3]          # BGR frames are converted on the device (OpenCV's fixed-point luma, bit-exact)
    if H % 8 or W % 8:
        raise ValueError(f"frame size {W}x{H} must be a multiple of 8 (three 2x2 pools, net.py:62,65,68)")
    if n == 0:
        return []
    ctx = deepc._ctx
    eng = ctx.engine(H, W, max_batch=n)
    while True:
        try:
            counts, offsets, kpts, refined = eng.infer_batch_host(frames, dust_bin_ids, refinenet is not None)
            break
        except N.CapacityError:
            eng = ctx.engine(H, W, max_batch=n, max_patches=2 * eng.max_patches)   # crowded frames: grow and retry
    return _rows_to_frames(counts, offsets, kpts, refined)


def infer_image(img: np.ndarray, dust_bin_ids: int, deepc: DeepcHandle,
                refinenet: Optional[RefineHandle] = None,
                draw_pred: bool = False,
                device='cpu'):
    """Do full inference on a BGR image -- inference.py:32-70, same arguments and return value.
    `device` is accepted for signature compatibility; the work always runs on the engine's B200."""
    # cv2.cvtColor(img, COLOR_BGR2GRAY) (:40) runs on the device inside the same launch sequence: OpenCV's fixed-point luma,
    # bit-exact with cv2 (tests/test_gpu_golden.py::test_bgr_batch_equals_gray_batch)
    assert img.ndim == 3 and img.shape[2] == 3, "infer_image expects a BGR image (H, W, 3)"
    frame = np.ascontiguousarray(img, dtype=np.uint8)[None]
    keypoints = infer_batch(frame, dust_bin_ids, deepc, refinenet)[0]
    if draw_pred:
        if refinenet is not None and keypoints.shape[0]:
            raw = infer_batch(frame, dust_bin_ids, deepc, None)[0]
            img = draw_inner_corners(img, raw[:, :2], raw[:, 2], radius=3, draw_ids=True, color=(0, 0, 255))
            img = draw_inner_corners(img, keypoints[:, :2], keypoints[:, 2], draw_ids=False, radius=1,
                                     color=(0, 255, 255))
        elif keypoints.shape[0]:
            img = draw_inner_corners(img, keypoints[:, :2], keypoints[:, 2], radius=3, draw_ids=True, color=(0, 0, 255))
    return keypoints, img


def pred_to_keypoints(loc_hat, ids_hat, dust_bin_ids: int):
    """model_utils.py:81-88 on the GPU: (N,65,h,w), (N,n_ids+1,h,w) logits -> (kpts (K,2) int64 [x,y], ids (K,) int64)
    in the reference's row-major order.  Like the reference, the batch index is dropped (:121-122)."""
    torch = _torch()
    assert loc_hat.ndim == 4 and ids_hat.ndim == 4                      # model_utils.py:85
    dev = loc_hat.device
    assert dev.type == "cuda", "pred_to_keypoints runs on the B200 engine; pass CUDA tensors"
    n, _, h, w = loc_hat.shape
    n_ids1 = int(ids_hat.shape[1])
    ctx = _scratch_context(n_ids1 - 1, dev.index or 0)
    cells = h * w
    eng = ctx.engine(h * 8, w * 8, max_batch=n, max_patches=n * cells)
    loc = loc_hat.to(torch.float32).contiguous()
    ids = ids_hat.to(torch.float32).contiguous()
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    offsets = torch.empty(n, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    kpts = torch.empty((n * cells, 4), dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    N.check(N.lib().dcu_decode_gather(eng.handle, loc.data_ptr(), ids.data_ptr(), None, n, int(dust_bin_ids), 0,
                                      counts.data_ptr(), offsets.data_ptr(), total.data_ptr(), kpts.data_ptr(), None, s))
    t = int(total.item())
    rows = kpts[:t].to(torch.int64)
    if t:
        frame = torch.repeat_interleave(torch.arange(n, device=dev), counts.to(torch.int64))
        order = torch.argsort(frame * cells + rows[:, 3], stable=True)   # (frame, cell) == torch.nonzero order
        rows = rows[order]
    return rows[:, :2].contiguous(), rows[:, 2].contiguous()


_SCRATCH = {}


class _DecodeContext:
    """Decode-only engines (no networks, no convolution workspace) for `pred_to_keypoints` / `extract_patches`, keyed by frame size."""

    def __init__(self, n_ids, device):
        self.n_ids, self.device = int(n_ids), int(device)
        self._engines = {}

    def engine(self, height, width, max_batch=1, max_patches=None) -> N.Engine:
        key = (int(height), int(width))
        eng = self._engines.get(key)
        want_p = max_patches or 256
        if eng is None or eng.max_batch < max_batch or eng.max_patches < want_p:
            if eng is not None:
                max_batch, want_p = max(max_batch, eng.max_batch), max(want_p, eng.max_patches)
                eng.close()
            eng = N.Engine(None, None, height, width, self.n_ids, self.device, max_batch=max_batch, max_patches=want_p,
                           decode_only=True)
            self._engines[key] = eng
        return eng


def _scratch_context(n_ids, device):
    key = (n_ids, device)
    if key not in _SCRATCH:
        N.lib()
        _SCRATCH[key] = _DecodeContext(n_ids, device)
    return _SCRATCH[key]


def extract_patches(img, keypoints, patch_size: int = 24):
    """model_utils.py:19-36 on the GPU: img (1,H,W) normalised float32 CUDA tensor, keypoints (K,2) int -> (K,24,24)."""
    torch = _torch()
    assert patch_size == 24, "the engine gathers the 24x24 patches RefineNet consumes"
    dev = img.device
    assert dev.type == "cuda", "extract_patches runs on the B200 engine; pass CUDA tensors"
    x = img.to(torch.float32).contiguous()
    H, W = int(x.shape[-2]), int(x.shape[-1])
    k = int(keypoints.shape[0])
    eng = _scratch_context(16, dev.index or 0).engine(H, W)
    xy = keypoints.to(dev).to(torch.int32).contiguous()
    out = torch.empty((k, 24, 24), dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    N.check(N.lib().dcu_extract_patches(eng.handle, x.data_ptr(), xy.data_ptr(), k, out.data_ptr(), s))
    return out


def resize_gpu(frames, input_size, device=0):
    """cv2.resize(frame, (W, H), interpolation=cv2.INTER_LINEAR) for a batch of uint8 frames (N,Hs,Ws) or (N,Hs,Ws,3) on the B200;
    bit-exact with cv2 when shrinking (the camera-to-network direction), DcuError (unsupported) when enlarging."""
    torch = _torch()
    frames = np.ascontiguousarray(frames, np.uint8)
    W, H = int(input_size[0]), int(input_size[1])
    n, Hs, Ws = frames.shape[:3]
    ch = 3 if frames.ndim == 4 else 1
    eng = _scratch_context(16, int(device)).engine(H, W)
    dev = torch.device("cuda", int(device))
    src = torch.from_numpy(frames).to(dev)
    dst = torch.empty((n, H, W, ch) if ch == 3 else (n, H, W), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev)
    N.check(N.lib().dcu_resize_u8(eng.handle, src.data_ptr(), n, Hs, Ws, ch, dst.data_ptr(), s.cuda_stream))
    s.synchronize()
    return dst.cpu().numpy()


def solve_pnp(keypoints, col_count, row_count, square_len, camera_matrix, dist_coeffs):
    """inference.py:15-29, unchanged in behaviour (host OpenCV on <= n_ids points; not a GPU target)."""
    import cv2
    if keypoints.shape[0] < 4:
        return False, None, None
    inn_rc = np.arange(1, row_count)
    inn_cc = np.arange(1, col_count)
    object_points = np.zeros(((col_count - 1) * (row_count - 1), 3), np.float32)
    object_points[:, :2] = np.array(np.meshgrid(inn_rc, inn_cc)).reshape((2, -1)).T * square_len
    image_points = keypoints[:, :2].astype(np.float32)
    object_points_found = object_points[keypoints[:, 2].astype(int)]
    ret, rvec, tvec = cv2.solvePnP(object_points_found, image_points, camera_matrix, dist_coeffs)
    return ret, rvec, tvec


def solve_pnp_batch(keypoints_list, col_count, row_count, square_len, camera_matrix, dist_coeffs, deepc: "DeepcHandle" = None):
    """`solve_pnp` for every frame of a batch in ONE GPU launch (the loop at pose_estimation.py:58-63).

    keypoints_list: what `infer_batch` returns (per frame a (K,3) array [x, y, id] or an empty array).  Returns a list of
    `(ret, rvec, tvec)` with the reference's conventions: `(False, None, None)` below 4 corners (inference.py:16-17), else
    `ret` bool and rvec / tvec as (3,1) float64 like cv2.solvePnP.  The solver restates cv2's SOLVEPNP_ITERATIVE for planar
    points in fp64 (csrc/pnp_core.cuh): same minimum as cv2 to ~1e-7 on well-conditioned frames; `solve_pnp` (host cv2)
    stays the drop-in default.  `deepc` selects the engine (device); any loaded model's engine is used when omitted."""
    ctx = deepc._ctx if deepc is not None else _any_context()
    eng = next(iter(ctx._engines.values())) if ctx._engines else ctx.engine(240, 320)
    n = len(keypoints_list)
    if n == 0:
        return []
    counts = np.array([0 if np.asarray(k).size == 0 else np.asarray(k).shape[0] for k in keypoints_list], np.int32)
    rows = [np.asarray(k, np.float64).reshape(-1, 3) for k in keypoints_list if np.asarray(k).size]
    flat = np.concatenate(rows, 0) if rows else np.zeros((0, 3))
    kpts = np.zeros((flat.shape[0], 4), np.int32)
    kpts[:, 2] = flat[:, 2].astype(int)                                         # keypoints[:, 2].astype(int), inference.py:26
    refined = flat[:, :2].astype(np.float32)                                    # keypoints[:, :2].astype(np.float32), :25
    ret, rvec, tvec = eng.solve_pnp_batch_host(counts, kpts, refined, col_count, row_count, square_len, camera_matrix, dist_coeffs)
    out = []
    for i in range(n):
        if counts[i] < 4:
            out.append((False, None, None))
        else:
            out.append((bool(ret[i]), rvec[i].reshape(3, 1).copy(), tvec[i].reshape(3, 1).copy()))
    return out


_CONTEXTS = []


def _any_context():
    live = [c for c in _CONTEXTS if c is not None]
    if not live:
        raise RuntimeError("solve_pnp_batch needs a loaded model (load_models) to pick a device")
    return live[-1]


def draw_inner_corners(img, corners, ids, radius=2, draw_ids=True, color=(0, 0, 255)):
    """Host-side drawing used only when draw_pred=True (reference: aruco_utils.py:135-192); copies the image."""
    import cv2
    out = img.copy()
    for (x, y), i in zip(np.asarray(corners), np.asarray(ids)):
        c = (int(round(float(x))), int(round(float(y))))
        cv2.circle(out, c, radius, color, -1)
        if draw_ids:
            cv2.putText(out, str(int(i)), (c[0] + 3, c[1] - 3), cv2.FONT_HERSHEY_SIMPLEX, 0.3, color, 1, cv2.LINE_AA)
    return out
