"""Detector validation metric at engine speed.

`DC_Metrics` mirrors /root/reference/src/models/metrics.py:38-146 (same constructor argument, `update(preds, target)`,
`compute()`, same accumulation), but the per-sample work -- decode of the logits, label decode, per-id worst distance and
match ratio -- runs in two kernels on the B200 (`dcu_decode_gather`, `dcu_dc_metrics`) instead of a Python loop over samples.
`update_frames` feeds uint8 frames through the engine's own detector first (what a validation loop over images does)."""
import ctypes as C

import numpy as np

from . import _native as N


class DC_Metrics:
    higher_is_better = False

    def __init__(self, dust_bin_ids, deepc):
        """deepc: a model handle from load_models (selects device and weights).  metrics.py:42-47."""
        self.dust_bin_ids = int(dust_bin_ids)
        self.px_margin = 3
        self._ctx = deepc._ctx
        self.distance = np.float32(0.0)
        self.ratio = np.float32(0.0)

    def _accumulate(self, l2, ratio, valid, bs):
        # metrics.py:56-73: sums over the samples that have labels, divided by the batch size
        if valid.any():
            self.distance = np.float32(self.distance + np.float32(l2[valid].sum(dtype=np.float32)) / np.float32(bs))
            self.ratio = np.float32(self.ratio + np.float32(ratio[valid].sum(dtype=np.float32)) / np.float32(bs))

    def _metrics(self, eng, counts, offsets, kpts, n, target):
        import torch
        loc_t, ids_t = target
        dev = counts.device
        loc_t = torch.as_tensor(loc_t).to(device=dev, dtype=torch.int64).contiguous()
        ids_t = torch.as_tensor(ids_t).to(device=dev, dtype=torch.int64).contiguous()
        assert loc_t.shape == ids_t.shape == (n, eng.height // 8, eng.width // 8), "labels must be (N, H/8, W/8)"
        l2 = torch.empty(n, dtype=torch.float32, device=dev)
        ratio = torch.empty(n, dtype=torch.float32, device=dev)
        valid = torch.empty(n, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().dcu_dc_metrics(eng.handle, counts.data_ptr(), offsets.data_ptr(), kpts.data_ptr(), n, loc_t.data_ptr(),
                                       ids_t.data_ptr(), self.dust_bin_ids, l2.data_ptr(), ratio.data_ptr(), valid.data_ptr(), s))
        l2, ratio, valid = l2.cpu().numpy(), ratio.cpu().numpy(), valid.cpu().numpy().astype(bool)
        self._accumulate(l2, ratio, valid, n)
        return l2, ratio, valid

    def update(self, preds, target):
        """preds = (loc_hat (N,65,h,w), ids_hat (N,n_ids+1,h,w)) float32 CUDA tensors, target = (loc (N,h,w), ids (N,h,w)) integer
        label maps -- metrics.py:49-51.  Returns the per-sample (l2, ratio, valid) arrays as a convenience."""
        import torch
        loc_x, ids_x = preds
        n, _, h, w = loc_x.shape
        # early-training / random logits keep almost every cell (the reference's own metrics.py self-test): capacity = every cell.
        # Decode + metric need no network: a decode-only engine (no conv workspace)
        from .inference import _scratch_context
        eng = _scratch_context(int(ids_x.shape[1]) - 1, self._ctx.device).engine(8 * h, 8 * w, max_batch=n, max_patches=n * h * w)
        dev = torch.device("cuda", eng.device)
        loc_x = loc_x.to(device=dev, dtype=torch.float32).contiguous()
        ids_x = ids_x.to(device=dev, dtype=torch.float32).contiguous()
        counts = torch.empty(n, dtype=torch.int32, device=dev)
        offsets = torch.empty(n, dtype=torch.int32, device=dev)
        total = torch.zeros(1, dtype=torch.int32, device=dev)
        kpts = torch.empty((eng.max_patches, 4), dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().dcu_decode_gather(eng.handle, loc_x.data_ptr(), ids_x.data_ptr(), None, n, self.dust_bin_ids, 0,
                                          counts.data_ptr(), offsets.data_ptr(), total.data_ptr(), kpts.data_ptr(), None, s))
        return self._metrics(eng, counts, offsets, kpts, n, target)

    def update_frames(self, frames_u8, target):
        """frames (N,H,W) uint8 -> the engine's detector + decode (no RefineNet) -> the same metric update."""
        import torch
        frames = np.ascontiguousarray(frames_u8, np.uint8)
        n, H, W = frames.shape
        eng = self._ctx.engine(H, W, max_batch=n)
        dev = torch.device("cuda", eng.device)
        fr = torch.from_numpy(frames).to(dev)
        while True:
            o = eng.infer_batch_device(fr.data_ptr(), n, self.dust_bin_ids, False, torch.cuda.current_stream(dev).cuda_stream)
            total = int(o["total"].item())
            if total <= eng.max_patches:
                break
            eng = self._ctx.engine(H, W, max_batch=n, max_patches=max(total, 2 * eng.max_patches))     # crowded frames: grow and retry
        return self._metrics(eng, o["counts"], o["offsets"], o["kpts"], n, target)

    def compute(self):
        """metrics.py:131-132."""
        return self.distance, self.ratio


class Refinenet_Metrics:
    """Mirror of /root/reference/src/models/metrics.py:135-161 with the per-sample arg-maxes and distances on the device
    (`dcu_refinenet_metrics`).  `update(preds, target)`: preds (N,1,64,64) or (N,64,64) heat maps, target (N,64,64) -- as the
    reference; `update_patches(patches, keypoints, target)`: runs the engine's own RefineNet on (N,24,24) patches and compares its
    arg-max with the target's."""
    higher_is_better = False

    def __init__(self, refinenet):
        self._ctx = refinenet._ctx
        self.distance = np.float32(0.0)

    def _run(self, eng, heat_pred, corners, target):
        import torch
        dev = torch.device("cuda", eng.device)
        target = torch.as_tensor(target).to(device=dev, dtype=torch.float32).contiguous()
        n = int(target.shape[0])
        assert tuple(target.shape[1:]) == (64, 64), "targets must be (N, 64, 64) heat maps"
        dist = torch.empty(n, dtype=torch.float32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().dcu_refinenet_metrics(eng.handle, None if heat_pred is None else heat_pred.data_ptr(),
                                              None if corners is None else corners.data_ptr(), target.data_ptr(), n, dist.data_ptr(), s))
        d = dist.cpu().numpy()
        if n:
            self.distance = np.float32(self.distance + d.mean(dtype=np.float32))       # metrics.py:157-158
        return d

    def update(self, preds, target):
        import torch
        eng = next(iter(self._ctx._engines.values())) if self._ctx._engines else self._ctx.engine(240, 320)
        dev = torch.device("cuda", eng.device)
        preds = torch.as_tensor(preds).to(device=dev, dtype=torch.float32)
        if preds.ndim == 4:
            preds = preds[:, 0]                                                        # loc_x.squeeze(1), metrics.py:143
        return self._run(eng, preds.contiguous(), None, target)

    def update_patches(self, patches, keypoints, target):
        import torch
        eng = next(iter(self._ctx._engines.values())) if self._ctx._engines else self._ctx.engine(240, 320)
        dev = torch.device("cuda", eng.device)
        patches = torch.as_tensor(patches).to(device=dev, dtype=torch.float32).reshape(-1, 24, 24).contiguous()
        n = int(patches.shape[0])
        if n > eng.max_patches:
            eng = self._ctx.engine(eng.height, eng.width, max_batch=eng.max_batch, max_patches=n)
        kp = torch.as_tensor(np.asarray(keypoints)).to(device=dev, dtype=torch.int32).reshape(-1, 2).contiguous()
        corners = torch.empty((n, 2), dtype=torch.int32, device=dev)
        refined = torch.empty((n, 2), dtype=torch.float32, device=dev)
        s = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().dcu_refine_forward(eng.handle, patches.data_ptr(), kp.data_ptr(), 2, n, corners.data_ptr(), refined.data_ptr(), None, s))
        return self._run(eng, None, corners, target)

    def compute(self):
        return self.distance


def _pack_targets(target_list):
    tc = np.array([0 if np.asarray(t).size == 0 else np.asarray(t).shape[0] for t in target_list], np.int32)
    rows = [np.asarray(t, np.float64).reshape(-1, 3) for t in target_list if np.asarray(t).size]
    flat = np.ascontiguousarray(np.concatenate(rows, 0) if rows else np.zeros((0, 3), np.float64))
    toff = np.concatenate([[0], np.cumsum(tc)[:-1]]).astype(np.int32) if len(tc) else np.zeros(0, np.int32)
    return tc, toff, flat


def pixel_error_batch(raw_list, refined_list, target_list, device=0):
    """`utils.pixel_error` (/root/reference/src/utils.py:33-52) for every frame of a batch in ONE kernel launch (`dcu_pixel_error`).

    raw_list / refined_list: what `infer_batch(..., None)` / `infer_batch(..., refinenet)` return (per frame (K,3) [x, y, id] or an
    empty array); target_list: per frame the label keypoints (T,3) [x, y, id] (float).  Returns (status int32 [N], out float64 [N,6]):
    out[f] = (mean raw, mean refined, mean refined-vs-raw, max raw, max refined, max refined-vs-raw) error in pixels, bit-identical
    with the reference's float64 numbers; status 1 = evaluated, 0 = skipped like the reference ((None, None) or no labels / corners),
    -1 = a frame the reference's numpy code would raise on."""
    import torch
    from .inference import _scratch_context
    n = len(raw_list)
    assert len(refined_list) == n and len(target_list) == n
    if n == 0:
        return np.zeros(0, np.int32), np.zeros((0, 6), np.float64)
    counts = np.array([0 if np.asarray(r).size == 0 else np.asarray(r).shape[0] for r in raw_list], np.int32)
    for r, f in zip(raw_list, refined_list):
        assert (np.asarray(r).size == 0) == (np.asarray(f).size == 0) and (np.asarray(r).size == 0 or np.asarray(r).shape == np.asarray(f).shape)
    total = int(counts.sum())
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    kpts = np.zeros((max(total, 1), 4), np.int32)
    refined = np.zeros((max(total, 1), 2), np.float32)
    if total:
        raw = np.concatenate([np.asarray(r).reshape(-1, 3) for r in raw_list if np.asarray(r).size], 0)
        ref = np.concatenate([np.asarray(r).reshape(-1, 3) for r in refined_list if np.asarray(r).size], 0)
        assert np.array_equal(raw[:, 2].astype(np.int64), ref[:, 2].astype(np.int64)), "raw and refined rows must list the same ids"
        kpts[:total, :3] = raw.astype(np.int64)
        refined[:total] = ref[:, :2]
    tc, toff, tflat = _pack_targets(target_list)
    eng = _scratch_context(16, int(device)).engine(240, 320, max_batch=1, max_patches=max(256, total))
    dev = torch.device("cuda", eng.device)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_counts, d_offsets, d_kpts, d_ref, d_tc, d_toff = d(counts), d(offsets), d(kpts), d(refined), d(tc), d(toff)
    d_t = d(tflat if tflat.size else np.zeros((1, 3)))
    status = torch.empty(n, dtype=torch.int32, device=dev)
    out = torch.empty((n, 6), dtype=torch.float64, device=dev)
    N.check(N.lib().dcu_pixel_error(eng.handle, d_counts.data_ptr(), d_offsets.data_ptr(), d_kpts.data_ptr(), d_ref.data_ptr(), n,
                                    d_tc.data_ptr(), d_toff.data_ptr(), d_t.data_ptr(), status.data_ptr(), out.data_ptr(),
                                    torch.cuda.current_stream(dev).cuda_stream))
    return status.cpu().numpy(), out.cpu().numpy()


def pixel_error(kpts_raw, kpts_ref, kpts_target, verbose=True, device=0):
    """Drop-in for `utils.pixel_error(kpts_raw, kpts_ref, kpts_target)` (utils.py:33-52): returns (mean raw error, mean refined
    error) in pixels, or (None, None) when a predicted id has no label; prints the reference's summary lines when `verbose`."""
    status, out = pixel_error_batch([kpts_raw], [kpts_ref], [kpts_target], device=device)
    if status[0] < 0:
        raise ValueError("operands could not be broadcast together (several predictions and several labels of one id)")
    if status[0] == 0:
        return None, None
    if verbose:
        found = np.unique(np.asarray(kpts_raw)[:, 2])
        print(f'Errors in pixels of the {len(found)}/{len(np.asarray(kpts_target)[:, 2])} kpts found:')
        print(f'Mean error raw: {out[0, 0]:<5.3f} Max error raw: {out[0, 3]:<5.3f}')
        print(f'Mean error ref: {out[0, 1]:<5.3f} Max error ref: {out[0, 4]:<5.3f}')
        print(f'Mean dist raw/ref: {out[0, 2]:<5.3f} Max dist raw/ref: {out[0, 5]:<5.3f}')
    return out[0, 0], out[0, 1]
