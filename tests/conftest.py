import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_sample():
    return load_golden("sample_image.npz")


@pytest.fixture(scope="session")
def golden_synth():
    return load_golden("synthetic_320x240_seed0.npz")


@pytest.fixture(scope="session")
def golden_edge():
    return load_golden("edge_cases.npz")


@pytest.fixture(scope="session")
def golden_640():
    return load_golden("synthetic_640x480_seed3.npz")


@pytest.fixture(scope="session")
def states():
    from deepcharuco_b200 import weights_io as W
    return W.load_state(W.DEFAULT_DEEPC), W.load_state(W.DEFAULT_REFINENET)


@pytest.fixture(scope="session")
def models():
    """(deepc, refinenet) handles of the CUDA engine -- GPU tests only."""
    import deepcharuco_b200 as dc
    return dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda")


@pytest.fixture(scope="session")
def models_ffma():
    """Same, pinned to the fp32 CUDA-core convolution kernels (DCU_CONV_FFMA)."""
    import deepcharuco_b200 as dc
    from deepcharuco_b200 import _native as N
    deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device="cuda")
    deepc._ctx.set_conv_impl(N.CONV_FFMA)
    return deepc, refinenet


def split_rows(rows, counts):
    out, o = [], 0
    for c in counts.tolist():
        out.append(rows[o:o + c])
        o += c
    return out
