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


def _gather_packed(counts, flat, world, rank, group, device):
    """All-gather per-rank (counts int64 [n_r], rows float64 [k_r, 3]) as TENSORS (NCCL moves CUDA tensors, gloo CPU tensors):
    sizes first, then the payloads padded to the largest shard.  ~24 B per corner: a few hundred kB for a 2048-frame batch."""
    import torch
    import torch.distributed as dist
    sizes = torch.tensor([counts.shape[0], flat.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    n_max, k_max = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    buf = torch.zeros(n_max + 3 * k_max, dtype=torch.float64, device=device)
    buf[:counts.shape[0]] = torch.from_numpy(counts.astype(np.float64)).to(device)
    if flat.shape[0]:
        buf[n_max:n_max + 3 * flat.shape[0]] = torch.from_numpy(np.ascontiguousarray(flat).reshape(-1)).to(device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    res = []
    for r in range(world):
        b = out[r].cpu().numpy()
        n_r, k_r = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        res.append((b[:n_r].astype(np.int64), b[n_max:n_max + 3 * k_r].reshape(k_r, 3)))
    return res


def infer_batch_distributed(frames, dust_bin_ids, deepc=None, refinenet=None, group=None, local_fn=None):
    """Multi-GPU `infer_batch` for one process per GPU (torchrun / torch.distributed): every rank passes the SAME (N,H,W)
    uint8 batch, runs its contiguous shard on its own device and engine, and the per-frame results (tiny: 24 B per corner) are
    all-gathered as packed tensors, so every rank returns the full list in frame order.  No collective touches the data path
    (SURVEY.md 8e); the gather moves results only.  `local_fn(frames_shard) -> list` replaces the engine call in CPU tests."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(frames), rank, world)
    if local_fn is None:
        from . import inference
        local_fn = lambda fr: inference.infer_batch(fr, dust_bin_ids, deepc, refinenet)
    mine = local_fn(frames[lo:hi]) if hi > lo else []
    integer = refinenet is None and deepc is not None
    if world == 1:
        return list(mine)
    backend = dist.get_backend(group)
    device = torch.device("cuda", deepc._ctx.device if deepc is not None else torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")
    counts, flat = pack_results(mine)
    gathered = _gather_packed(counts, flat, world, rank, group, device)
    return merge_shards([unpack_results(c, f, integer=integer) for c, f in gathered])
