"""``deconv(lbda=None)``: noise-constrained regularisation (pybold/bold_signal.py:99-214).

Row N1 of SURVEY.md 8(f).  The inner prox-gradient loops run in the persistent ``pb_deconv``
kernel (warm-started through ``w0``); the outer loop -- one scalar update of alpha / lambda per
voxel and per outer iteration -- is driven from the host with O(V) tensor arithmetic.

The noise level sigma is the MAD of the level-1 db3 detail coefficients in the reference
(pybold/utils.py:10-25, PyWavelets); here one kernel (``pb_mad_daub_noise_est_*``).  The loop itself
is pinned on the live reference with sigma injected (tests/golden/deconv_auto.npz); the wavelet
convention is pinned on PyWavelets' documented Haar examples only (PyWavelets is not installed where
this was built) -- pass ``sigma=`` to ``deconv`` to bypass it.
"""
from __future__ import annotations

import torch

from . import _lib
from ._array import like_input, pick_dtype, ptr, stream_ptr, to_device

# the analysis high-pass the kernel applies (csrc/pb_noise.cuh holds the same six numbers)
_DB3_DEC_HI = (-0.3326705529509569, 0.8068915093133388, -0.4598775021193313,
               -0.13501102001039084, 0.08544127388224149, 0.035226291882100656)


_warned_sigma = False


def _mad_rows(x, c, entry):
    dtype = pick_dtype(x)
    xd = to_device(x, dtype)
    one_d = xd.dim() <= 1
    x2 = xd.reshape(1, -1) if one_d else xd.reshape(xd.shape[0], -1)
    V, n = x2.shape
    if n == 0:
        raise ValueError("%s: empty series" % entry)
    out = torch.empty(V, dtype=dtype, device=x2.device)
    with torch.cuda.device(x2.device):
        rc = _lib.fn(entry, dtype)(ptr(x2), float(c), ptr(out), V, n, stream_ptr())
    _lib.check(rc, entry)
    if one_d:
        return float(out[0])
    return like_input(out, x)


def mad(x, c=0.6744):
    """Median absolute deviation ``median(|x - median(x)|) / c`` (pybold/utils.py:10-13).

    A 1-D array (NumPy or torch) gives a float like the reference; a ``[V, n]`` batch gives one
    value per row.  Exact order statistics on the device (``pb_mad_*``).
    """
    return _mad_rows(x, c, "pb_mad")


def mad_daub_noise_est(x, c=0.6744):
    """Noise level from the MAD of the level-1 db3 detail coefficients (pybold/utils.py:16-25),
    1-D -> float, ``[V, T]`` -> one sigma per voxel (``pb_mad_daub_noise_est_*``).

    The wavelet step follows PyWavelets' ``wavedec(x, 'db3', level=1)`` convention (symmetric
    extension); PyWavelets is not installed where this was built, so that convention is pinned on
    PyWavelets' documented Haar examples and on the defining properties of the db3 filter only
    (tests/test_boundary.py), not on a PyWavelets db3 output.
    """
    return _mad_rows(x, c, "pb_mad_daub_noise_est")


def deconv_auto_lbda(y_in, yb, one_d, hrf, lipschitz, sigma, early_stopping, tol, wind,
                     nb_iter, nb_sub_iter):
    """Outer loop of ``deconv(lbda=None)`` (pybold/bold_signal.py:99-214) for a batch.

    Every outer iteration is two launches: the inner prox-gradient loops of the voxels that are still
    running (``pb_deconv_masked``: warm start, no cost trace, finished voxels skipped) and the fused
    residual / alpha / lambda update (``pb_noise_step``).  The alpha-window stop of every voxel is
    evaluated on the device; the host only looks at "is anything still running" every few iterations.
    """
    from .bold_signal import deconv_batch
    V, T = yb.shape
    dtype, dev = yb.dtype, yb.device
    if sigma is None:
        # bold_signal.py:103.  The wavelet convention of this estimate could not be compared with PyWavelets
        # itself (see mad_daub_noise_est): say so once, `sigma=` makes the run independent of it
        global _warned_sigma
        if not _warned_sigma:
            import warnings
            warnings.warn("pybold_b200.deconv(lbda=None): sigma comes from the on-device db3 MAD estimate, whose "
                          "boundary convention is pinned on PyWavelets' documented examples only; pass sigma= for "
                          "a run that does not depend on it", UserWarning, stacklevel=3)
            _warned_sigma = True
        sigma = mad_daub_noise_est(yb)
    sigma = torch.as_tensor(sigma, dtype=dtype, device=dev).reshape(-1).expand(V).clone()
    alpha = torch.ones(V, dtype=dtype, device=dev)                  # bold_signal.py:104
    lbda = 1.0 / (2.0 * alpha)
    mu = 1.0e-4
    w = torch.zeros_like(yb)
    x = torch.zeros_like(yb)
    z = torch.zeros_like(yb)
    scratch = (torch.empty_like(yb), torch.empty_like(yb), torch.empty_like(yb),
               torch.zeros(V, dtype=torch.int32, device=dev))
    active = torch.ones(V, dtype=torch.uint8, device=dev)
    n_outer = torch.full((V,), nb_iter, dtype=torch.int64, device=dev)   # outer iterations each voxel runs
    sub = int(wind / 2)
    hist = []
    J = torch.full((V, nb_iter), float("nan"), dtype=dtype, device=dev)
    R = torch.full((V, nb_iter), float("nan"), dtype=dtype, device=dev)
    G = torch.full((V, nb_iter), float("nan"), dtype=dtype, device=dev)
    step = _lib.fn("pb_noise_step", dtype)
    check_every = 1 if V == 1 else 4
    for i in range(nb_iter):
        x_n, z_n, w_n, _, _ = deconv_batch(yb, hrf, lbda, lipschitz, w, early_stopping, tol, wind,
                                           nb_sub_iter, active=active, out=scratch, trace=False)
        r = torch.empty(V, dtype=dtype, device=dev)
        g = torch.empty(V, dtype=dtype, device=dev)
        # bold_signal.py:139-145: keep the result of the active voxels, r, g, alpha and lambda updates
        with torch.cuda.device(dev):
            rc = step(ptr(x_n), ptr(z_n), ptr(w_n), ptr(yb), ptr(sigma), ptr(active), mu, ptr(x), ptr(z), ptr(w),
                      ptr(alpha), ptr(lbda), ptr(r), ptr(g), V, T, stream_ptr())
        _lib.check(rc, "pb_noise_step")
        hist.append(alpha.clone())
        if len(hist) > wind:
            hist = hist[1:]
        live = active.bool()
        R[:, i] = torch.where(live, r, R[:, i])
        G[:, i] = torch.where(live, g, G[:, i])
        J[:, i] = torch.where(live, 0.5 * r + lbda * g, J[:, i])
        if early_stopping and i > wind and sub > 0:                 # bold_signal.py:164-178
            old_it = torch.stack(hist[:-sub]).mean(dim=0)
            new_it = torch.stack(hist[-sub:]).mean(dim=0)
            stop = ((new_it - old_it).abs() / new_it.abs() < tol) & live
            n_outer = torch.where(stop, torch.full_like(n_outer, i + 1), n_outer)
            active = active & (~stop).to(torch.uint8)
            if (i % check_every == 0 or i == nb_iter - 1) and not bool(active.any()):
                break
    # bold_signal.py:180-212: last deconvolution with the final lambda, every voxel
    x, z, w, _, _ = deconv_batch(yb, hrf, lbda, lipschitz, w, early_stopping, tol, wind, nb_sub_iter, trace=False)
    if one_d:
        n = int(n_outer[0])
        conv = lambda t: like_input(t[0], y_in)  # noqa: E731
        return (conv(x), conv(z), conv(w), [float(v) for v in J[0, :n]], [float(v) for v in R[0, :n]],
                [float(v) for v in G[0, :n]])
    n_max = int(n_outer.max())
    conv = lambda t: like_input(t, y_in)  # noqa: E731
    return conv(x), conv(z), conv(w), conv(J[:, :n_max]), conv(R[:, :n_max]), conv(G[:, :n_max])
