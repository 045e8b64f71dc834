"""One process per GPU (torchrun): `infer_batch_distributed` under the NCCL backend.  Every rank passes the same batch, runs its
shard on its own B200 and receives the full result list; rank 0 checks it against its own single-GPU run of the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepcharuco_b200 as dc  # noqa: E402
from deepcharuco_b200 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device=local)
frames = synth.make_frames(41, 240, 320, seed=13)          # odd count: shards differ in size
ok = True
for ref in (refinenet, None):
    full = dc.infer_batch_distributed(frames, 16, deepc, ref)
    assert len(full) == len(frames)
    if rank == 0:
        want = dc.infer_batch(frames, 16, deepc, ref)
        for a, b in zip(full, want):
            ok = ok and a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
if rank == 0:
    print("DIST_CHECK", "OK" if int(flag.item()) == 1 else "MISMATCH", "world", world, "corners", sum(0 if r.size == 0 else r.shape[0] for r in full))
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
