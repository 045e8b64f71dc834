"""CPU oracle for the DeepCharuco hot path -- TEST INFRASTRUCTURE ONLY.

A restatement of the reference algorithm (JunkyByte/deepcharuco @ 37d569fc)
for the path  detector forward -> decode + patch gather -> RefineNet forward
+ sub-pixel argmax.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import this package; the product
package `deepcharuco_b200` never does (tests/test_layout.py enforces it).

Arithmetic note: the reference's conv / batch-norm / pool / argmax arithmetic
lives in a third-party dependency that is not under /root/reference --
PyTorch (pinned torch==2.1.0 in the reference's src/requirements.txt; this
image has torch 2.11.0, CPU = ATen + oneDNN).  The oracle therefore calls the
same torch CPU functional ops in fp32, in the reference's order; the integer
decode is restated in numpy.

Parity pin: the reference ships no tests and no golden vectors (SURVEY.md
section 4).  The oracle is pinned instead against outputs of the reference
itself, run in the build container by tools/make_golden.py and committed as
tests/golden/*.npz (tests/test_oracle_golden.py), and -- when /root/reference
is mounted -- directly against the live reference (tests/test_oracle_vs_reference.py).
"""
from .nets import detector_forward, refinenet_forward, load_state  # noqa: F401
from .decode import (pre_bgr_image, pred_argmax, label_to_keypoints,  # noqa: F401
                     pred_to_keypoints, extract_patches, bargmax2d,
                     refine_corners, marshal_keypoints)
from .pipeline import infer_image, infer_gray_batch, solve_pnp  # noqa: F401
from . import metrics  # noqa: F401
