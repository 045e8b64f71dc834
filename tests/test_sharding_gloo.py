"""N > 1 host path on CPU: two gloo ranks shard a batch, process their shard, gather, and the merged result equals
the single-process result.  The per-frame work is stood in by the oracle's decode on golden logits (no GPU here);
what is under test is shard_range / pack / gather / merge -- the only cross-rank logic the engine has."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frame_result(loc, ids):
    sys.path.insert(0, ROOT)
    import oracle
    k, i = oracle.pred_to_keypoints(loc[None], ids[None], 16)
    return oracle.marshal_keypoints(k, i) if len(i) else np.array([])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from deepcharuco_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "synthetic_320x240_seed0.npz"))
    loc, ids = z["loc"], z["ids"]
    n = 5                                            # cycle the 3 golden frames to an odd count
    lo, hi = sharding.shard_range(n, rank, world)
    mine = [_frame_result(loc[i % 3], ids[i % 3]) for i in range(lo, hi)]
    counts, flat = sharding.pack_results(mine)
    gathered = [None] * world
    dist.all_gather_object(gathered, (counts, flat))
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the timing reduction bench.py does (max over ranks)
    # the user-facing wrapper: same batch on every rank, full result list on every rank
    frames = np.arange(n, dtype=np.int64)          # stand-in "frames": indices into the golden logits
    full = sharding.infer_batch_distributed(frames, 16, local_fn=lambda fr: [_frame_result(loc[i % 3], ids[i % 3]) for i in fr])
    assert len(full) == n
    for i, got in enumerate(full):
        want = _frame_result(loc[i % 3], ids[i % 3])
        assert got.shape == want.shape and np.array_equal(np.asarray(got, np.float64), np.asarray(want, np.float64))
    # fewer frames than ranks (rank 1's shard is empty) and a frame without corners
    one = sharding.infer_batch_distributed(frames[:1], 16, local_fn=lambda fr: [_frame_result(loc[i % 3], ids[i % 3]) for i in fr])
    assert len(one) == 1 and np.array_equal(np.asarray(one[0], np.float64), np.asarray(_frame_result(loc[0], ids[0]), np.float64))
    none = sharding.infer_batch_distributed(frames[:3], 16, local_fn=lambda fr: [np.array([]) for _ in fr])
    assert len(none) == 3 and all(r.size == 0 for r in none)
    if rank == 0:
        merged = sharding.merge_shards([sharding.unpack_results(c, f, integer=True) for c, f in gathered])
        np.savez(os.path.join(out_dir, "merged.npz"), n=len(merged), tmax=t.numpy(),
                 **{f"r{i}": m for i, m in enumerate(merged)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_merge(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "merged.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "synthetic_320x240_seed0.npz"))
    assert int(z["n"]) == 5 and float(z["tmax"][0]) == 2.0
    for i in range(5):
        want = _frame_result(g["loc"][i % 3], g["ids"][i % 3])
        assert np.array_equal(z[f"r{i}"], want)
