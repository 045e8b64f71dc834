"""Batch sharding across the GPUs of one box: frames are independent, so the batch is split into
contiguous ranges, one per rank, with NO collective on the data path (replicas, batch-sharded).
"""
import numpy as np


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous [start, stop) of `n_items` owned by `rank`; sizes differ by at most one."""
    assert 0 <= rank < world_size
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_shards(per_rank_results):
    """Concatenate per-rank lists of per-frame keypoint arrays back into frame order."""
    out = []
    for r in per_rank_results:
        out.extend(r)
    return out


def pack_results(results):
    """list of per-frame (K,3) arrays -> (counts int64 [N], rows float64 [sum K, 3]) for a gather over ranks."""
    counts = np.array([0 if r.size == 0 else r.shape[0] for r in results], np.int64)
    rows = [np.asarray(r, np.float64).reshape(-1, 3) for r in results if r.size]
    flat = np.concatenate(rows, 0) if rows else np.zeros((0, 3), np.float64)
    return counts, flat


def unpack_results(counts, flat, integer=False):
    out, o = [], 0
    for c in counts.tolist():
        if c == 0:
            out.append(np.array([]))
        else:
            r = flat[o:o + c]
            out.append(r.astype(np.int64) if integer else r.copy())
            o += c
    return out
