"""``deconv(lbda=None)``: noise-constrained regularisation (pybold/bold_signal.py:99-214).

Row N1 of SURVEY.md 8(f).  The inner prox-gradient loops run in the persistent ``pb_deconv``
kernel (warm-started through ``w0``); the outer loop -- one scalar update of alpha / lambda per
voxel and per outer iteration -- is driven from the host with O(V) tensor arithmetic.

The noise level sigma is the MAD of the level-1 db3 detail coefficients in the reference
(pybold/utils.py:10-25, PyWavelets).  PyWavelets is not available where this was built, so
``mad_daub_noise_est`` below follows pywt's documented convention but is NOT pinned against
it; pass ``sigma=`` to ``deconv`` for a reference-exact run.
"""
from __future__ import annotations

import torch

from . import _lib
from ._array import like_input, ptr, stream_ptr

_DB3_DEC_HI = (-0.3326705529509569, 0.8068915093133388, -0.4598775021193313,
               -0.13501102001039084, 0.08544127388224149, 0.035226291882100656)


def mad_daub_noise_est(y, c=0.6744):
    """sigma[v] = median(|cD - median(cD)|) / c on a ``[V, T]`` CUDA tensor (utils.py:10-25)."""
    V, T = y.shape
    F = len(_DB3_DEC_HI)
    ext = torch.cat([y[:, :F - 1].flip(1), y, y[:, -(F - 1):].flip(1)], dim=1)
    n_out = (T + F - 1) // 2
    taps = torch.tensor(_DB3_DEC_HI, dtype=y.dtype, device=y.device)
    idx = 2 * torch.arange(n_out, device=y.device) + 1 + (F - 1)
    cD = torch.zeros((V, n_out), dtype=y.dtype, device=y.device)
    for j in range(F):
        cD += taps[j] * ext[:, idx - j]
    med = cD.median(dim=1, keepdim=True).values if n_out % 2 else _median(cD)
    dev = (cD - med).abs()
    mad = dev.median(dim=1, keepdim=True).values if n_out % 2 else _median(dev)
    return (mad / c).reshape(-1)


def mad(x, c=0.6744):
    """Median absolute deviation of each row (pybold/utils.py:10-13); 1-D input gives a scalar tensor."""
    x2 = x.reshape(1, -1) if x.dim() == 1 else x
    n = x2.shape[1]
    med = x2.median(dim=1, keepdim=True).values if n % 2 else _median(x2)
    dev = (x2 - med).abs()
    out = (dev.median(dim=1, keepdim=True).values if n % 2 else _median(dev)) / c
    return out.reshape(()) if x.dim() == 1 else out.reshape(-1)


def _median(a):
    """NumPy-style median (mean of the two middle values for an even count)."""
    s, _ = torch.sort(a, dim=1)
    n = a.shape[1]
    return 0.5 * (s[:, n // 2 - 1:n // 2] + s[:, n // 2:n // 2 + 1])


def deconv_auto_lbda(y_in, yb, one_d, hrf, lipschitz, sigma, early_stopping, tol, wind,
                     nb_iter, nb_sub_iter):
    from .bold_signal import deconv_batch
    V, T = yb.shape
    dtype, dev = yb.dtype, yb.device
    if sigma is None:
        sigma = mad_daub_noise_est(yb)
    sigma = torch.as_tensor(sigma, dtype=dtype, device=dev).reshape(-1).expand(V).clone()
    alpha = torch.ones(V, dtype=dtype, device=dev)                  # bold_signal.py:104
    lbda = 1.0 / (2.0 * alpha)
    mu = 1.0e-4
    w = torch.zeros_like(yb)
    x = torch.zeros_like(yb)
    z = torch.zeros_like(yb)
    active = torch.ones(V, dtype=torch.uint8, device=dev)
    sub = int(wind / 2)
    hist = []
    J, R, G = [], [], []
    step = _lib.fn("pb_noise_step", dtype)
    for i in range(nb_iter):
        x_n, z_n, w_n, _, _ = deconv_batch(yb, hrf, lbda, lipschitz, w, early_stopping, tol, wind,
                                           nb_sub_iter)
        r = torch.empty(V, dtype=dtype, device=dev)
        g = torch.empty(V, dtype=dtype, device=dev)
        # bold_signal.py:139-145: keep the result of the active voxels, r, g, alpha and lambda updates
        rc = step(ptr(x_n), ptr(z_n), ptr(w_n), ptr(yb), ptr(sigma), ptr(active), mu, ptr(x), ptr(z), ptr(w),
                  ptr(alpha), ptr(lbda), ptr(r), ptr(g), V, T, stream_ptr())
        _lib.check(rc, "pb_noise_step")
        hist.append(alpha.clone())
        if len(hist) > wind:
            hist = hist[1:]
        R.append(r)
        G.append(g)
        J.append(0.5 * r + lbda * g)
        if early_stopping and i > wind and sub > 0:                 # bold_signal.py:164-178
            old_it = torch.stack(hist[:-sub]).mean(dim=0)
            new_it = torch.stack(hist[-sub:]).mean(dim=0)
            stop = (new_it - old_it).abs() / new_it.abs() < tol
            active = active & (~stop).to(torch.uint8)
            if not bool(active.any()):
                break
    # bold_signal.py:180-212: last deconvolution with the final lambda
    x, z, w, _, _ = deconv_batch(yb, hrf, lbda, lipschitz, w, early_stopping, tol, wind, nb_sub_iter)
    J, R, G = torch.stack(J, 1), torch.stack(R, 1), torch.stack(G, 1)
    if one_d:
        conv = lambda t: like_input(t[0], y_in)  # noqa: E731
        return (conv(x), conv(z), conv(w), [float(v) for v in J[0]], [float(v) for v in R[0]],
                [float(v) for v in G[0]])
    conv = lambda t: like_input(t, y_in)  # noqa: E731
    return conv(x), conv(z), conv(w), conv(J), conv(R), conv(G)
