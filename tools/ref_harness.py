"""Import the UNMODIFIED reference (JunkyByte/deepcharuco) from /root/reference.

Test/fixture tooling only: used by tools/make_golden.py and by the `not gpu`
tests that cross-check `oracle/` against the real reference when the mount is
present (this container).  It never runs on the GPU box (no /root/reference
there) and nothing in the product package imports it.

The reference imports `pytorch_lightning` and `torchmetrics` at module load
(src/models/net.py:6, src/models/refinenet.py:6, src/models/metrics.py:3);
neither is installed here, so two minimal stand-ins are registered first.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DEEPCHARUCO_REF", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
DEEPC_CKPT = os.path.join(REF_SRC, "reference", "longrun-epoch=99-step=369700.ckpt")
REFINE_CKPT = os.path.join(REF_SRC, "reference", "second-refinenet-epoch-100-step=373k.ckpt")
SAMPLE_IMAGE = os.path.join(REF_SRC, "reference", "samples_test", "IMG_7412.png")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "inference.py"))


def _install_stubs():
    import torch
    from torch import nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            @classmethod
            def load_from_checkpoint(cls, path, **kw):
                obj = cls(**kw)
                sd = torch.load(path, map_location="cpu", weights_only=True)["state_dict"]
                obj.load_state_dict(sd, strict=True)
                return obj

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl

    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def add_state(self, name, default, dist_reduce_fx=None):
                setattr(self, name, default)

        tm.Metric = Metric
        sys.modules["torchmetrics"] = tm


_ref = None


def load_reference():
    """Returns the reference's `inference` module (imported once)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError(f"reference not mounted at {REF_ROOT}")
    _install_stubs()
    for p in (REF_SRC, os.path.join(REF_SRC, "models")):
        if p not in sys.path:
            sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir(REF_SRC)  # the reference resolves paths relative to src/
    try:
        import inference as ref_inference  # noqa
    finally:
        os.chdir(cwd)
    _ref = ref_inference
    return _ref


def load_reference_models(device="cpu"):
    ref = load_reference()
    return ref.load_models(DEEPC_CKPT, REFINE_CKPT, n_ids=16, device=device)
