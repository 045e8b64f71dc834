"""Two engines on two GPUs of ONE process (the in-process alternative to one process per GPU): same results on both, and the
batch split / merged by sharding.shard_range equals the single-device result.  Skipped on a one-GPU box."""
import numpy as np
import pytest
import torch

import deepcharuco_b200 as dc
from deepcharuco_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def test_two_devices_in_one_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    frames = synth.make_frames(24, 240, 320, seed=13)
    m0 = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda:0")
    m1 = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda:1")
    a = dc.infer_batch(frames, 16, *m0)
    b = dc.infer_batch(frames, 16, *m1)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    parts = []
    for rank, m in enumerate((m0, m1)):
        lo, hi = sharding.shard_range(len(frames), rank, 2)
        parts.append(dc.infer_batch(frames[lo:hi], 16, *m))
    merged = sharding.merge_shards(parts)
    assert len(merged) == len(a) and all(np.array_equal(x, y) for x, y in zip(merged, a))
    one = dc.infer_batch(frames[:1], 16, *m1)[0]          # small-batch graph path on the second device
    one = dc.infer_batch(frames[:1], 16, *m1)[0]
    one = dc.infer_batch(frames[:1], 16, *m1)[0]
    assert np.array_equal(one, a[0])


def test_infer_batch_distributed_under_nccl():
    """One process per GPU (torchrun, NCCL backend): the user-facing multi-GPU call returns, on every rank, exactly what one GPU
    returns for the whole batch (tools/dist_check.py).  Needs two GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + (os.getpid() % 300)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "dist_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-3000:]
