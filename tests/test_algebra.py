"""CPU checks of the two algebraic rewrites the CUDA kernels rely on (pure torch / numpy, no GPU):

* UP mode (conv_tc2.cu, engine.cu: collapse_up_weights): a 3x3 same-padded convolution of a 2x nearest-upsampled map equals, per
  output phase (a, b), a 2x2 convolution of the LOW-resolution map with the kernel rows / columns that land on the same pixel summed
  (refinenet.py:66-67, 71-72, 76-77 are the layer pairs this replaces).
* FLAT mode (conv_tc2.cu): a valid 3x3 convolution of a stack of small maps equals a 1-D "shifted run" product over the
  concatenated pixels, tap (ky, kx) = shift by ky*row + kx, with the wrap-around positions dropped; with one zero gutter column / row
  per map the same holds for same-padded convolutions (the gutter is the padding of every neighbour)."""
import numpy as np
import torch
import torch.nn.functional as F


def test_upsample_then_conv_equals_phase_collapsed_2x2():
    g = torch.Generator().manual_seed(0)
    L = torch.randn(2, 5, 7, 6, generator=g, dtype=torch.float64)
    w = torch.randn(4, 5, 3, 3, generator=g, dtype=torch.float64)
    want = F.conv2d(F.interpolate(L, scale_factor=2, mode="nearest"), w, padding=1)
    lo = [[0, 1], [0, 2]]          # [phase][k] -> first kernel index that lands on low-res offset k
    hi = [[0, 2], [1, 2]]          # ... last
    Lp = F.pad(L, (1, 1, 1, 1))
    got = torch.zeros_like(want)
    for a in range(2):
        for b in range(2):
            acc = 0
            for ky in range(2):
                for kx in range(2):
                    wc = w[:, :, lo[a][ky]:hi[a][ky] + 1, lo[b][kx]:hi[b][kx] + 1].sum((2, 3))         # [o, i]
                    # low-res pixel (y + a - 1 + ky, x + b - 1 + kx)
                    sl = Lp[:, :, a + ky:a + ky + L.shape[2], b + kx:b + kx + L.shape[3]]
                    acc = acc + torch.einsum("oi,nihw->nohw", wc, sl)
            got[:, :, a::2, b::2] = acc
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def _flat_conv(run, w, row, pad):
    """run: [C, P] pixel run; returns [O, P] with out[j] = sum_taps w[:, :, ky, kx] @ run[j + (ky-pad)*row + (kx-pad)] (zeros outside)."""
    C, P = run.shape
    out = np.zeros((w.shape[0], P))
    for ky in range(3):
        for kx in range(3):
            off = (ky - pad) * row + (kx - pad)
            sh = np.zeros_like(run)
            if off >= 0:
                sh[:, :P - off] = run[:, off:]
            else:
                sh[:, -off:] = run[:, :P + off]
            out += w[:, :, ky, kx] @ sh
    return out


def test_flat_run_equals_valid_conv():
    rng = np.random.default_rng(1)
    n, C, O, H, W = 5, 3, 4, 6, 7
    x = rng.standard_normal((n, C, H, W))
    w = rng.standard_normal((O, C, 3, 3))
    want = F.conv2d(torch.from_numpy(x), torch.from_numpy(w)).numpy()                     # valid: (H-2) x (W-2)
    run = x.transpose(1, 0, 2, 3).reshape(C, n * H * W)
    out = _flat_conv(run, w, W, 0).reshape(O, n, H, W).transpose(1, 0, 2, 3)
    assert np.allclose(out[:, :, :H - 2, :W - 2], want, rtol=1e-12, atol=1e-12)           # wrap-around positions are dropped


def test_flat_run_with_gutters_equals_same_conv():
    rng = np.random.default_rng(2)
    n, C, O, S = 6, 3, 4, 8
    x = rng.standard_normal((n, C, S, S))
    w = rng.standard_normal((O, C, 3, 3))
    want = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy()
    cell = np.zeros((n, C, S + 1, S + 1))                                                  # 8x8 data in a 9x9 cell, zero gutter column / row
    cell[:, :, :S, :S] = x
    run = cell.transpose(1, 0, 2, 3).reshape(C, n * (S + 1) * (S + 1))
    out = _flat_conv(run, w, S + 1, 1).reshape(O, n, S + 1, S + 1).transpose(1, 0, 2, 3)
    assert np.allclose(out[:, :, :S, :S], want, rtol=1e-12, atol=1e-12)
