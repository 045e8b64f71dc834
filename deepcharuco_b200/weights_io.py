"""Weight loading: Lightning .ckpt (as shipped by the reference) or the converted .npz.

Reference loader replaced: inference.load_models -> lModel.load_from_checkpoint
(/root/reference/src/inference.py:73-84).  The checkpoints are dicts whose
'state_dict' keys are 'model.<layer>.{weight,bias,running_mean,running_var,
num_batches_tracked}' (SURVEY.md 3.2); they load with weights_only=True.
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_DEEPC = os.path.join(_HERE, "weights", "deepc.npz")
DEFAULT_REFINENET = os.path.join(_HERE, "weights", "refinenet.npz")


def load_state(path):
    """-> {'<layer>.<param>': float32 ndarray} (no 'model.' prefix, no num_batches_tracked)."""
    path = str(path)
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files}
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=True)["state_dict"]
    out = {}
    for k, v in sd.items():
        if not k.startswith("model.") or k.endswith("num_batches_tracked"):
            continue
        out[k[len("model."):]] = np.ascontiguousarray(v.detach().cpu().numpy(), dtype=np.float32)
    return out


def save_state(state, path):
    np.savez(path, **state)
