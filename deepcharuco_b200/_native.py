"""ctypes binding of libdeepcharuco_b200.so (C ABI: include/deepcharuco_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C deepcharuco_b200/csrc`.
There is no fallback: if the shared object is missing or the device is not sm_100 the
import / engine creation raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCU_LIB_PATH") or os.path.join(_HERE, "libdeepcharuco_b200.so")   # env override: A/B builds

DCU_OK, DCU_ERR_INVALID, DCU_ERR_CUDA, DCU_ERR_CAPACITY, DCU_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
CONV_FFMA, CONV_TCGEN05 = 0, 1
FLAG_DECODE_ONLY = 1
CONV_DEFAULT = CONV_TCGEN05   # tcgen05 tensor-core path (fp16 hi/lo split); CONV_FFMA is the strict-fp32 CUDA-core path

EXPORTS = [
    "dcu_create", "dcu_destroy", "dcu_detector_forward", "dcu_detector_forward_f32", "dcu_extract_patches",
    "dcu_decode_gather", "dcu_refine_forward", "dcu_infer_batch", "dcu_infer_batch_host", "dcu_infer_batch_host_bgr", "dcu_bgr_to_gray",
    "dcu_debug_conv_layer", "dcu_debug_tc_stats", "dcu_set_conv_impl", "dcu_launch_count", "dcu_profile_enable", "dcu_profile_read", "dcu_profile_read_issued", "dcu_solve_pnp_batch", "dcu_solve_pnp_batch_host", "dcu_dc_metrics", "dcu_refinenet_metrics", "dcu_pixel_error", "dcu_synth_frames", "dcu_warp_perspective_u8", "dcu_resize_u8", "dcu_profile_records", "dcu_detector_flops_per_frame",
    "dcu_refine_flops_per_patch", "dcu_last_error", "dcu_version",
]


class DcuConvLayer(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p),
                ("cin", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32)]


class DcuSynthFrame(C.Structure):
    _fields_ = [("lat_step", C.c_int32), ("lat_h", C.c_int32), ("lat_w", C.c_int32), ("n_boards", C.c_int32),
                ("bg_lo", C.c_float), ("bg_hi", C.c_float), ("gain", C.c_float), ("reserved", C.c_float),
                ("blur_w", C.c_float * 16), ("hinv", (C.c_double * 9) * 4)]


class DcuConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("n_ids", C.c_int32),
                ("max_batch", C.c_int32), ("max_patches", C.c_int32), ("conv_impl", C.c_int32),
                ("reserved", C.c_int32)]


class DcuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"deepcharuco_b200 error {code}: {msg}")
        self.code = code


class CapacityError(DcuError):
    pass


_lib = None


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C deepcharuco_b200/csrc`); there is no CPU or PyTorch fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.dcu_create.argtypes = [C.POINTER(DcuConfig), C.POINTER(DcuConvLayer), i32, C.POINTER(DcuConvLayer), i32, C.POINTER(vp)]
    L.dcu_destroy.argtypes = [vp]
    L.dcu_detector_forward.argtypes = [vp, vp, i32, vp, vp, vp]
    L.dcu_detector_forward_f32.argtypes = [vp, vp, i32, vp, vp, vp]
    L.dcu_extract_patches.argtypes = [vp, vp, vp, i32, vp, vp]
    L.dcu_decode_gather.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.dcu_refine_forward.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.dcu_infer_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.dcu_infer_batch_host.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.dcu_infer_batch_host_bgr.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.dcu_bgr_to_gray.argtypes = [vp, vp, i32, vp, vp]
    L.dcu_resize_u8.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.dcu_debug_conv_layer.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp]
    L.dcu_debug_tc_stats.argtypes = [vp, i32, vp]
    L.dcu_set_conv_impl.argtypes = [vp, i32]
    L.dcu_profile_enable.argtypes = [vp, i32]
    L.dcu_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    L.dcu_profile_read_issued.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.dcu_refinenet_metrics.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    L.dcu_synth_frames.argtypes = [vp, vp, i32, C.c_uint64, i32, vp, i32, vp, vp]
    L.dcu_warp_perspective_u8.argtypes = [vp, vp, i32, i32, vp, vp, i32, i32, vp]
    L.dcu_pixel_error.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.dcu_dc_metrics.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, vp]
    L.dcu_solve_pnp_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, C.c_double, vp, vp, i32, vp, vp, vp, vp]
    L.dcu_solve_pnp_batch_host.argtypes = [vp, vp, vp, vp, i32, i32, i32, C.c_double, vp, vp, i32, vp, vp, vp, vp]
    L.dcu_profile_records.argtypes = [vp, i32, vp, C.POINTER(i32)]
    L.dcu_launch_count.argtypes = [vp]
    L.dcu_launch_count.restype = i64
    L.dcu_detector_flops_per_frame.argtypes = [vp]
    L.dcu_detector_flops_per_frame.restype = C.c_double
    L.dcu_refine_flops_per_patch.argtypes = [vp]
    L.dcu_refine_flops_per_patch.restype = C.c_double
    L.dcu_last_error.restype = C.c_char_p
    L.dcu_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc):
    if rc == DCU_OK:
        return
    msg = lib().dcu_last_error().decode("utf-8", "replace")
    if rc == DCU_ERR_CAPACITY:
        raise CapacityError(rc, msg)
    raise DcuError(rc, msg)


# ---------------------------------------------------------------------------------------------------
# layer tables (the order dcu_create expects) and BatchNorm folding
# ---------------------------------------------------------------------------------------------------
DET_LAYERS = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
              "convPa", "convPb", "convDa", "convDb"]          # net.py:22-48
REF_LAYERS = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
              "conv5a", "conv5b", "convPa", "convPb"]          # refinenet.py:22-47
BN_EPS = np.float32(1e-5)


def fold_bn(state, name):
    """Eval-mode BatchNorm2d as y = fma(x, alpha, beta) with alpha = gamma * fl32(1/sqrt(var+eps)) and
    beta = bn_bias - mean*alpha rounded once.  Bit-identical to ATen's CPU batch_norm (checked in
    tests/test_host_logic.py); do not fold into the conv weights."""
    bn = "bn" + name[len("conv"):]
    if bn + ".weight" not in state:
        return None, None
    g, b = state[bn + ".weight"], state[bn + ".bias"]
    m, v = state[bn + ".running_mean"], state[bn + ".running_var"]
    inv = (np.float32(1) / np.sqrt(v + BN_EPS)).astype(np.float32)
    alpha = (g * inv).astype(np.float32)
    beta = (b.astype(np.float64) - m.astype(np.float64) * alpha.astype(np.float64)).astype(np.float32)
    return np.ascontiguousarray(alpha), np.ascontiguousarray(beta)


def layer_table(state, names):
    """-> (ctypes array of DcuConvLayer, keep-alive list of ndarrays)."""
    arr = (DcuConvLayer * len(names))()
    keep = []
    for i, n in enumerate(names):
        w = np.ascontiguousarray(state[n + ".weight"], dtype=np.float32)
        b = np.ascontiguousarray(state[n + ".bias"], dtype=np.float32)
        a, e = fold_bn(state, n)
        keep += [w, b, a, e]
        arr[i].weight = w.ctypes.data
        arr[i].bias = b.ctypes.data
        arr[i].alpha = a.ctypes.data if a is not None else None
        arr[i].beta = e.ctypes.data if e is not None else None
        arr[i].cout, arr[i].cin, arr[i].ksize = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
    return arr, keep


class Engine:
    """One C engine: fixed (device, H, W, n_ids), workspace for max_batch frames / max_patches corners."""

    def __init__(self, state_det, state_ref, height, width, n_ids=16, device=0, max_batch=1, max_patches=None,
                 conv_impl=CONV_DEFAULT, decode_only=False):
        L = lib()
        if max_patches is None:
            max_patches = max(256, 64 * max_batch)
        self.cfg = DcuConfig(device=int(device), height=int(height), width=int(width), n_ids=int(n_ids),
                             max_batch=int(max_batch), max_patches=int(max_patches), conv_impl=int(conv_impl),
                             reserved=FLAG_DECODE_ONLY if decode_only else 0)
        self.decode_only = bool(decode_only)
        if decode_only:          # no networks, no conv workspace: dcu_decode_gather / dcu_extract_patches on caller-owned buffers
            det, keep_d, n_det = None, [], 0
            state_ref = None
        else:
            det, keep_d = layer_table(state_det, DET_LAYERS)
            n_det = len(DET_LAYERS)
        if state_ref is not None:
            ref, keep_r = layer_table(state_ref, REF_LAYERS)
            n_ref = len(REF_LAYERS)
        else:
            ref, keep_r, n_ref = None, [], 0
        h = C.c_void_p()
        check(L.dcu_create(C.byref(self.cfg), det, n_det, ref, n_ref, C.byref(h)))
        self._h = h
        self.height, self.width, self.n_ids = int(height), int(width), int(n_ids)
        self.max_batch, self.max_patches = int(max_batch), int(max_patches)
        self.device = int(device)
        self.has_ref = state_ref is not None
        self.conv_impl = int(conv_impl)
        # host result buffers reused across calls
        self._counts = np.empty(self.max_batch, np.int32)
        self._offsets = np.empty(self.max_batch, np.int32)
        self._kpts = np.empty((self.max_patches, 4), np.int32)
        self._refined = np.empty((self.max_patches, 2), np.float32)
        self._total = C.c_int32(0)

    def close(self):
        if getattr(self, "_h", None):
            lib().dcu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_conv_impl(self, impl):
        check(lib().dcu_set_conv_impl(self._h, int(impl)))
        self.conv_impl = int(impl)

    def launch_count(self):
        return int(lib().dcu_launch_count(self._h))

    def profile_enable(self, on=True):
        check(lib().dcu_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, cls=0):
        ms, work, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        check(lib().dcu_profile_read(self._h, int(cls), C.byref(ms), C.byref(work), C.byref(n)))
        return ms.value, work.value, n.value

    def profile_read_issued(self, cls=0):
        w = C.c_double(0)
        check(lib().dcu_profile_read_issued(self._h, int(cls), C.byref(w)))
        return w.value

    def profile_records(self, cap=4096):
        """Per-launch records since profile_enable: float64 [n][8] = cls, ms, work, cin, cout, hout, wout, n."""
        rec = np.zeros((cap, 8), np.float64)
        n = C.c_int32(0)
        check(lib().dcu_profile_records(self._h, cap, rec.ctypes.data, C.byref(n)))
        return rec[:min(cap, n.value)]

    def infer_batch_device(self, frames_dev_ptr, n, dust_bin_ids=16, use_refinenet=True, stream=None):
        """Frames already resident in HBM (device pointer); results stay in the engine's device buffers.
        Returns nothing: used by bench.py's device-resident leg, which only needs the work to happen."""
        import torch
        if not hasattr(self, "_dev_out"):
            dev = torch.device("cuda", self.device)
            self._dev_out = dict(counts=torch.empty(self.max_batch, dtype=torch.int32, device=dev),
                                 offsets=torch.empty(self.max_batch, dtype=torch.int32, device=dev),
                                 total=torch.zeros(1, dtype=torch.int32, device=dev),
                                 kpts=torch.empty((self.max_patches, 4), dtype=torch.int32, device=dev),
                                 refined=torch.empty((self.max_patches, 2), dtype=torch.float32, device=dev))
        o = self._dev_out
        check(lib().dcu_infer_batch(self._h, frames_dev_ptr, int(n), int(dust_bin_ids), 1 if use_refinenet else 0,
                                    o["counts"].data_ptr(), o["offsets"].data_ptr(), o["total"].data_ptr(),
                                    o["kpts"].data_ptr(), o["refined"].data_ptr(), stream))
        return o

    def solve_pnp_batch_host(self, counts, kpts, refined, col_count, row_count, square_len, camera_matrix, dist_coeffs, stream=None):
        """counts int32 [N], kpts int32 [sum,4] (x, y, id, cell), refined float32 [sum,2] or None -> (ret int32 [N], rvec f64 [N,3], tvec f64 [N,3])."""
        counts = np.ascontiguousarray(counts, np.int32)
        kpts = np.ascontiguousarray(kpts, np.int32).reshape(-1, 4)
        n = int(counts.shape[0])
        assert int(counts.sum()) == kpts.shape[0]
        ref = None if refined is None else np.ascontiguousarray(refined, np.float32).reshape(-1, 2)
        cam = np.ascontiguousarray(camera_matrix, np.float64).reshape(3, 3)
        dist = np.zeros(0) if dist_coeffs is None else np.ascontiguousarray(dist_coeffs, np.float64).reshape(-1)
        ret, rvec, tvec = np.zeros(n, np.int32), np.zeros((n, 3), np.float64), np.zeros((n, 3), np.float64)
        check(lib().dcu_solve_pnp_batch_host(self._h, counts.ctypes.data, kpts.ctypes.data, None if ref is None else ref.ctypes.data, n,
                                             int(col_count), int(row_count), float(square_len), cam.ctypes.data,
                                             dist.ctypes.data if dist.size else None, int(dist.size), ret.ctypes.data,
                                             rvec.ctypes.data, tvec.ctypes.data, stream))
        return ret, rvec, tvec

    def solve_pnp_batch_device(self, n, col_count, row_count, square_len, camera_matrix, dist_coeffs, use_refined=True, stream=None):
        """Pose of every frame of the last infer_batch_device call, on its device-resident results; returns torch tensors
        (ret int32 [n], rvec f64 [n,3], tvec f64 [n,3]) on the engine's device."""
        import torch
        o = self._dev_out
        dev = o["counts"].device
        if "pnp_ret" not in o or o["pnp_ret"].shape[0] < n:
            o["pnp_ret"] = torch.zeros(self.max_batch, dtype=torch.int32, device=dev)
            o["pnp_rvec"] = torch.zeros((self.max_batch, 3), dtype=torch.float64, device=dev)
            o["pnp_tvec"] = torch.zeros((self.max_batch, 3), dtype=torch.float64, device=dev)
        cam = np.ascontiguousarray(camera_matrix, np.float64).reshape(3, 3)
        dist = np.zeros(0) if dist_coeffs is None else np.ascontiguousarray(dist_coeffs, np.float64).reshape(-1)
        check(lib().dcu_solve_pnp_batch(self._h, o["counts"].data_ptr(), o["offsets"].data_ptr(), o["kpts"].data_ptr(),
                                        o["refined"].data_ptr() if use_refined else None, int(n), int(col_count), int(row_count),
                                        float(square_len), cam.ctypes.data, dist.ctypes.data if dist.size else None, int(dist.size),
                                        o["pnp_ret"].data_ptr(), o["pnp_rvec"].data_ptr(), o["pnp_tvec"].data_ptr(), stream))
        return o["pnp_ret"][:n], o["pnp_rvec"][:n], o["pnp_tvec"][:n]

    def detector_flops_per_frame(self):
        return float(lib().dcu_detector_flops_per_frame(self._h))

    def refine_flops_per_patch(self):
        return float(lib().dcu_refine_flops_per_patch(self._h))

    def infer_batch_host(self, frames_u8, dust_bin_ids=16, use_refinenet=True, stream=None):
        """frames_u8: (N,H,W) grayscale or (N,H,W,3) BGR uint8 host array -> (counts[N], offsets[N], kpts[total,4] int32, refined[total,2] f32 | None).
        H2D, the whole pipeline and D2H happen inside the call (dcu_infer_batch_host)."""
        f = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        n = int(f.shape[0])
        bgr = f.ndim == 4 and f.shape[3] == 3
        if f.ndim not in (3, 4) or (f.ndim == 4 and not bgr) or f.shape[1] != self.height or f.shape[2] != self.width:
            raise ValueError(f"frames must be (N,{self.height},{self.width}) or (N,{self.height},{self.width},3) uint8, got {f.shape}")
        entry = lib().dcu_infer_batch_host_bgr if bgr else lib().dcu_infer_batch_host
        rc = entry(self._h, f.ctypes.data, n, int(dust_bin_ids), 1 if use_refinenet else 0,
                                        self._counts.ctypes.data, self._offsets.ctypes.data, C.addressof(self._total),
                                        self._kpts.ctypes.data, self._refined.ctypes.data if use_refinenet else None,
                                        stream)
        check(rc)
        total = int(self._total.value)
        return (self._counts[:n].copy(), self._offsets[:n].copy(), self._kpts[:total].copy(),
                self._refined[:total].copy() if use_refinenet else None)
