"""Convert the reference's two Lightning checkpoints to plain fp32 .npz state files.

    python tools/convert_ckpt.py            # /root/reference/src/reference/*.ckpt -> deepcharuco_b200/weights/*.npz

The .npz files are committed: the GPU box has no /root/reference, and parity on
trained weights (near-tied argmax margins) is the point of the tests.
The checkpoints are MIT-licensed data shipped in the reference repo (README.md:89-107).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.dirname(__file__))
import ref_harness as rh  # noqa: E402
from deepcharuco_b200 import weights_io as W  # noqa: E402

if __name__ == "__main__":
    for src, dst in ((rh.DEEPC_CKPT, W.DEFAULT_DEEPC), (rh.REFINE_CKPT, W.DEFAULT_REFINENET)):
        st = W.load_state(src)
        W.save_state(st, dst)
        n = sum(v.size for v in st.values())
        print(f"{src} -> {dst}: {len(st)} tensors, {n} floats")
