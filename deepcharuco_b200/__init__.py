"""deepcharuco_b200 -- B200-native (sm_100a) engine for the DeepCharuco inference hot path.

Public surface mirrors /root/reference/src/inference.py; see inference.py in this package.
"""
from .inference import (load_models, infer_image, infer_batch, solve_pnp, solve_pnp_batch, pred_to_keypoints,  # noqa: F401
                        extract_patches, pre_bgr_image)
from .sharding import infer_batch_distributed  # noqa: F401
from .weights_io import DEFAULT_DEEPC, DEFAULT_REFINENET  # noqa: F401
