"""Oracle restatement of the reference's validation metric for the detector (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/models/metrics.py: DC_Metrics (:38-146) on top of pred_to_keypoints / label_to_keypoints
(:18-35, same decode as model_utils.py).  Distances are evaluated in float32 like torch.cdist on float32 keypoints; all
coordinates are integers, so every distance is a correctly rounded sqrt of an exact integer."""
import numpy as np

from .decode import label_to_keypoints, pred_to_keypoints

PX_MARGIN = 3            # metrics.py:46


def _per_id_max_dist(keypoints, ids, target_keypoints, target_ids):
    """metrics.py:83-99 / :110-127 -- for every unique target id (ascending, torch.unique): the largest distance between the
    predictions with that id and the label with that id; ids without a prediction are skipped."""
    out = []
    for i, idv in enumerate(np.unique(target_ids)):
        mask = np.nonzero(ids == idv)[0]
        tmask = np.nonzero(target_ids == idv)[0]
        if mask.size == 0 or tmask.size == 0:
            continue
        a = keypoints[mask].astype(np.float32)[:, None, :]
        b = target_keypoints[tmask].astype(np.float32)[None, :, :]
        d = np.sqrt(((a - b) ** 2).sum(-1, dtype=np.float32)).astype(np.float32)     # torch.cdist(p=2)
        out.append((i, np.float32(d.max())))
    return out


def compute_l2_distance(keypoints, ids, target_keypoints, target_ids):
    """metrics.py:102-129 -- sum of the per-id worst distances / max(1, ids found); None without labels."""
    if len(target_ids) == 0:
        return None
    per = _per_id_max_dist(keypoints, ids, target_keypoints, target_ids)
    distances = np.zeros(len(target_ids), np.float32)
    for i, d in per:
        distances[i] = d
    return np.float32(distances.sum(dtype=np.float32) / np.float32(max(1, len(per))))


def compute_ratio(keypoints, ids, target_keypoints, target_ids):
    """metrics.py:75-100 -- fraction of labels whose id was predicted within PX_MARGIN pixels (worst prediction counts)."""
    if len(target_ids) == 0:
        return None
    matches = np.zeros(len(target_ids), np.float32)
    for i, d in _per_id_max_dist(keypoints, ids, target_keypoints, target_ids):
        if d < PX_MARGIN:
            matches[i] = 1
    return np.float32(matches.mean(dtype=np.float32))


class DCMetrics:
    """metrics.py:38-73,131-132 -- state (distance, ratio); update() adds the batch means over ALL samples of the batch."""

    def __init__(self, dust_bin_ids):
        self.dust_bin_ids = dust_bin_ids
        self.distance = np.float32(0)
        self.ratio = np.float32(0)

    def sample(self, loc_x, ids_x, loc_target, ids_target):
        kp, idv = pred_to_keypoints(loc_x[None], ids_x[None], self.dust_bin_ids)
        kt, it = label_to_keypoints(loc_target[None], ids_target[None], self.dust_bin_ids)
        return compute_l2_distance(kp, idv, kt, it), compute_ratio(kp, idv, kt, it)

    def update(self, preds, target):
        (loc_x, ids_x), (loc_target, ids_target) = preds, target
        bs = loc_x.shape[0]
        l2_sum, ratio_sum, atleast = np.float32(0), np.float32(0), False
        for i in range(bs):
            l2, ratio = self.sample(loc_x[i], ids_x[i], loc_target[i], ids_target[i])
            if l2 is not None:
                atleast = True
                l2_sum = np.float32(l2_sum + l2)
                ratio_sum = np.float32(ratio_sum + ratio)
        if atleast:
            self.distance = np.float32(self.distance + l2_sum / np.float32(bs))
            self.ratio = np.float32(self.ratio + ratio_sum / np.float32(bs))

    def compute(self):
        return self.distance, self.ratio


class RefinenetMetrics:
    """metrics.py:135-161 -- per update: mean over the batch of the L2 distance (in heat-map pixels) between the arg-max of the
    predicted 64x64 heat map and the arg-max of the target map (first maximum of the flattened map, like torch.argmax)."""

    def __init__(self):
        self.distance = np.float32(0)

    @staticmethod
    def per_sample(preds, target):
        p = np.asarray(preds, np.float32).reshape(len(preds), -1)
        t = np.asarray(target, np.float32).reshape(len(target), -1)
        d = np.int64(np.asarray(target).shape[-1])
        mp, mt = p.argmax(1), t.argmax(1)
        a = np.stack((mp // d, mp % d), 1).astype(np.float32)
        b = np.stack((mt // d, mt % d), 1).astype(np.float32)
        return np.sqrt(((a - b) ** 2).sum(1, dtype=np.float32)).astype(np.float32)

    def update(self, preds, target):
        self.distance = np.float32(self.distance + self.per_sample(preds, target).mean(dtype=np.float32))

    def compute(self):
        return self.distance


def utils_l2_distance(keypoints, ids, target_keypoints, target_ids):
    """/root/reference/src/utils.py:6-30 -- distances = zeros(len(target_ids)); for i, id in enumerate(unique(target_ids)): the
    largest float64 distance between the keypoints with that id and the target(s) with that id (numpy broadcasting, raises like
    the reference on (m,2) - (t,2) with m != t, m != 1, t != 1); ids nobody predicted keep 0.  None without targets."""
    distances = np.zeros((len(target_ids),))
    if distances.size == 0:
        return None
    for i, idv in enumerate(np.unique(target_ids)):
        mask = np.nonzero(ids == idv)[0]
        tmask = np.nonzero(target_ids == idv)[0]
        if mask.size == 0 or tmask.size == 0:
            continue
        diff = keypoints[mask] - target_keypoints[tmask]
        dist = np.sqrt((diff * diff).sum(axis=1))            # np.linalg.norm(ord=2, axis=1) on real input
        distances[i] = np.max(dist)
    return distances


def pixel_error(kpts_raw, kpts_ref, kpts_target):
    """utils.py:33-52 without the prints -> (status, [mean d, mean d_ref, mean d_raw_ref, max d, max d_ref, max d_raw_ref]);
    status 0 = the reference returns (None, None) (a raw id without a label)."""
    if not set(kpts_raw[:, 2]).issubset(set(kpts_target[:, 2])):
        return 0, None
    d = utils_l2_distance(kpts_raw[:, :2], kpts_raw[:, 2], kpts_target[:, :2], kpts_target[:, 2])
    d_ref = utils_l2_distance(kpts_ref[:, :2], kpts_ref[:, 2], kpts_target[:, :2], kpts_target[:, 2])
    d_rr = utils_l2_distance(kpts_ref[:, :2], kpts_ref[:, 2], kpts_raw[:, :2], kpts_raw[:, 2])
    return 1, np.array([d.mean(), d_ref.mean(), d_rr.mean(), d.max(), d_ref.max(), d_rr.max()])
