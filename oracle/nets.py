"""Oracle restatement of the two networks (torch CPU fp32 functional ops).

Follows, layer for layer:
  * dcModel.forward        -- /root/reference/src/models/net.py:50-80
    (layer table net.py:12-48: 3x3 s1 p1 conv + BatchNorm2d(eps=1e-5) + ReLU,
    MaxPool2d(2,2) after 1b/2b/3b, two heads 3x3 128->256 then 1x1 -> 65 / n_ids+1)
  * RefineNet.forward      -- /root/reference/src/models/refinenet.py:49-83
    (layer table refinenet.py:10-47: four valid 3x3 convs, pool, 3a/3b, up x2,
    4a/4b, up x2, 5a/5b, up x2, Pa, 1x1 Pb)
Weights come from a plain {name: ndarray} state (see load_state); no Lightning.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, never overridden by the reference


def load_state(path):
    """Load a `{layer.param: float32 ndarray}` state from .npz or a Lightning .ckpt.

    The .ckpt layout: dict with 'state_dict' whose keys are
    'model.<layer>.<param>' (inference.py:74,80 load it through Lightning).
    """
    if str(path).endswith(".npz"):
        with np.load(path) as z:
            return {k: np.ascontiguousarray(z[k]) for k in z.files}
    sd = torch.load(path, map_location="cpu", weights_only=True)["state_dict"]
    out = {}
    for k, v in sd.items():
        if not k.startswith("model.") or k.endswith("num_batches_tracked"):
            continue
        out[k[len("model."):]] = v.detach().cpu().numpy().astype(np.float32, copy=False)
    return out


def _t(state, name):
    return torch.from_numpy(np.ascontiguousarray(state[name]))


def _conv_bn_relu(x, state, name, pad):
    """conv(+bias) -> BatchNorm2d(eval, running stats) -> ReLU (net.py:60, refinenet.py:56)."""
    bn = "bn" + name[len("conv"):]
    x = F.conv2d(x, _t(state, name + ".weight"), _t(state, name + ".bias"), stride=1, padding=pad)
    x = F.batch_norm(x, _t(state, bn + ".running_mean"), _t(state, bn + ".running_var"),
                     _t(state, bn + ".weight"), _t(state, bn + ".bias"),
                     training=False, momentum=0.1, eps=BN_EPS)
    return F.relu(x)


@torch.no_grad()
def detector_forward(state, x, return_features=False):
    """x: (N,1,H,W) float32 tensor -> loc (N,65,H/8,W/8), ids (N,n_ids+1,H/8,W/8).  net.py:50-80."""
    feats = {}
    x = _conv_bn_relu(x, state, "conv1a", 1); feats["conv1a"] = x
    x = _conv_bn_relu(x, state, "conv1b", 1)
    x = F.max_pool2d(x, 2, 2); feats["conv1b"] = x          # net.py:62 (indices unused)
    x = _conv_bn_relu(x, state, "conv2a", 1); feats["conv2a"] = x
    x = _conv_bn_relu(x, state, "conv2b", 1)
    x = F.max_pool2d(x, 2, 2); feats["conv2b"] = x          # net.py:65
    x = _conv_bn_relu(x, state, "conv3a", 1); feats["conv3a"] = x
    x = _conv_bn_relu(x, state, "conv3b", 1)
    x = F.max_pool2d(x, 2, 2); feats["conv3b"] = x          # net.py:68
    x = _conv_bn_relu(x, state, "conv4a", 1); feats["conv4a"] = x
    x = _conv_bn_relu(x, state, "conv4b", 1); feats["conv4b"] = x
    cPa = _conv_bn_relu(x, state, "convPa", 1); feats["convPa"] = cPa
    loc = F.conv2d(cPa, _t(state, "convPb.weight"), _t(state, "convPb.bias"))   # net.py:74, no activation
    cDa = _conv_bn_relu(x, state, "convDa", 1); feats["convDa"] = cDa
    ids = F.conv2d(cDa, _t(state, "convDb.weight"), _t(state, "convDb.bias"))   # net.py:77
    if return_features:
        return loc, ids, feats
    return loc, ids


@torch.no_grad()
def refinenet_forward(state, x, return_features=False):
    """x: (K,1,24,24) float32 tensor -> heat (K,1,64,64).  refinenet.py:49-83."""
    feats = {}
    x = _conv_bn_relu(x, state, "conv1a", 0); feats["conv1a"] = x   # 22x22
    x = _conv_bn_relu(x, state, "conv1b", 0); feats["conv1b"] = x   # 20x20
    x = _conv_bn_relu(x, state, "conv2a", 0); feats["conv2a"] = x   # 18x18
    x = _conv_bn_relu(x, state, "conv2b", 0)                        # 16x16
    x = F.max_pool2d(x, 2, 2); feats["conv2b"] = x                  # 8x8   refinenet.py:62
    x = _conv_bn_relu(x, state, "conv3a", 1); feats["conv3a"] = x
    x = _conv_bn_relu(x, state, "conv3b", 1)
    x = F.interpolate(x, scale_factor=2, mode="nearest"); feats["conv3b"] = x   # UpsamplingNearest2d, :67
    x = _conv_bn_relu(x, state, "conv4a", 1); feats["conv4a"] = x
    x = _conv_bn_relu(x, state, "conv4b", 1)
    x = F.interpolate(x, scale_factor=2, mode="nearest"); feats["conv4b"] = x   # :72
    x = _conv_bn_relu(x, state, "conv5a", 1); feats["conv5a"] = x
    x = _conv_bn_relu(x, state, "conv5b", 1)
    x = F.interpolate(x, scale_factor=2, mode="nearest"); feats["conv5b"] = x   # :77
    cPa = _conv_bn_relu(x, state, "convPa", 1); feats["convPa"] = cPa
    heat = F.conv2d(cPa, _t(state, "convPb.weight"), _t(state, "convPb.bias"))  # :81
    if return_features:
        return heat, feats
    return heat
