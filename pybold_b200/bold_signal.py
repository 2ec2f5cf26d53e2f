"""Batched device solvers behind the signatures of ``pybold/bold_signal.py``.

``deconv`` and ``bd`` keep the reference's arguments, defaults and return tuples for a 1-D
voxel; a ``[V, T]`` input solves V voxels in ONE persistent kernel launch (the reference
fans out one process per voxel with joblib, examples/icassp_2019/validation.py:43-47).
NumPy in -> NumPy out (float64 unless everything passed is float32); torch CUDA tensors in ->
torch tensors out with no host round trip.

Deliberate, documented differences from the reference:
 * nothing is printed per iteration (the reference prints unconditionally, bold_signal.py:79-80);
 * the theta step is an exact bounded 1-D minimisation instead of SciPy's finite-difference
   L-BFGS-B (bold_signal.py:329-333): the two agree to ~1e-7 on theta (parity);
 * ``deconv`` needs the power-iteration start: it is drawn from ``np.random`` exactly like the
   reference (utils.py:97) unless ``x0=`` is passed.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._array import like_input, per_voxel, pick_dtype, ptr, stream_ptr, to_device, upload
from .hrf_model import MAX_DELTA, MIN_DELTA, hrf_len
from .linear import ConvAndLinear, DiscretInteg
from .utils import spectral_radius_est


def _as_batch(y, dtype):
    yd = to_device(y, dtype)
    if yd.dim() == 1:
        return yd.reshape(1, -1), True
    if yd.dim() != 2:
        raise ValueError("y must be 1-D (one voxel) or [V, T]")
    return yd, False


def deconv_batch(y, hrf, lbda, lipschitz, w0=None, early_stopping=True, tol=1.0e-6, wind=6,
                 nb_iter=1000, events=None, active=None, out=None, trace=True):
    """Device entry point: tensors in, tensors out.  Returns (x, z, diff_z, J_raw, n_iter).

    ``J_raw[v, k]`` is the un-normalised cost of iteration k (NaN past ``n_iter[v]``);
    ``lipschitz`` is the constant actually used (0.9 x power estimate in ``deconv``).
    ``events=(e0, e1)``: CUDA events recorded immediately around the solver launch (benchmarks).
    ``active`` (uint8 ``[V]``): voxels with 0 are skipped, their rows of the outputs stay as they are;
    ``out=(x, z, diff_z, n_iter)`` reuses output tensors; ``trace=False`` skips the cost trace (J is None).
    """
    V, T = y.shape
    dtype, dev = y.dtype, y.device
    hrf = hrf.to(device=dev, dtype=dtype).contiguous()
    K = hrf.shape[-1]
    if hrf.dim() > 2 or (hrf.dim() == 2 and hrf.shape[0] not in (1, V)):
        raise ValueError("hrf must be [K] (shared) or [V, K] with V = %d voxels, got %s"
                         % (V, tuple(hrf.shape)))
    h_stride = 0 if hrf.dim() == 1 or hrf.shape[0] == 1 else K
    if w0 is not None:
        if tuple(w0.shape) != (V, T) or w0.dtype != dtype or w0.device != dev:
            raise ValueError("w0 must be a %s tensor of shape (%d, %d) on %s" % (dtype, V, T, dev))
        w0 = w0.contiguous()
    y = y.contiguous()
    lb, lb_stride = per_voxel(lbda, V, dtype, dev, "lbda")
    Lc, L_stride = per_voxel(lipschitz, V, dtype, dev, "lipschitz")
    if out is not None:
        x, z, dz, n_iter = out
        for t in (x, z, dz):
            if tuple(t.shape) != (V, T) or t.dtype != dtype or t.device != dev or not t.is_contiguous():
                raise ValueError("out tensors must be contiguous %s tensors of shape (%d, %d) on %s" % (dtype, V, T, dev))
    else:
        x = torch.empty_like(y)
        z = torch.empty_like(y)
        dz = torch.empty_like(y)
        n_iter = torch.zeros(V, dtype=torch.int32, device=dev)
    J = torch.full((V, nb_iter), float("nan"), dtype=dtype, device=dev) if trace else None
    if active is not None and (active.dtype != torch.uint8 or active.numel() != V or active.device != dev):
        raise ValueError("active must be a uint8 tensor with one entry per voxel on %s" % dev)
    with torch.cuda.device(dev):        # the C side sizes its grid for, and launches on, the current device
        if events is not None:
            events[0].record()
        if active is None and trace:
            rc = _lib.fn("pb_deconv", dtype)(
                ptr(y), ptr(hrf), h_stride, ptr(Lc), L_stride, ptr(lb), lb_stride, ptr(w0),
                int(nb_iter), int(bool(early_stopping)), int(wind), float(tol),
                ptr(x), ptr(z), ptr(dz), ptr(J), ptr(n_iter), V, T, K, stream_ptr())
        else:
            rc = _lib.fn("pb_deconv_masked", dtype)(
                ptr(y), ptr(hrf), h_stride, ptr(Lc), L_stride, ptr(lb), lb_stride, ptr(w0), ptr(active),
                int(nb_iter), int(bool(early_stopping)), int(wind), float(tol),
                ptr(x), ptr(z), ptr(dz), ptr(J), ptr(n_iter), V, T, K, stream_ptr())
        if events is not None:
            events[1].record()
    _lib.check(rc, "pb_deconv")
    return x, z, dz, J, n_iter


def _lipschitz_cst(est, dtype):
    """``0.9 * spectral_radius_est`` (bold_signal.py:52), the product taken in double like the reference does
    and rounded once to the build's type: the same constant on every path that solves the same problem."""
    if isinstance(est, torch.Tensor):
        return (0.9 * est.double()).to(dtype)
    est = np.asarray(est, dtype=np.float64)
    return 0.9 * float(est) if est.ndim == 0 else 0.9 * est       # scalars stay host floats (cached on the device)


def deconv(y, t_r, hrf, lbda=None, early_stopping=True, tol=1.0e-6,  # noqa
           wind=6, nb_iter=1000, nb_sub_iter=1000, verbose=0, x0=None, sigma=None, dtype=None):
    """Sparse deconvolution with a known HRF (pybold/bold_signal.py:13-214).

    Returns ``(x, z, diff_z, J, R, G)`` like the reference: with a float ``lbda`` J is the cost
    trace normalised by its first entry and R, G are None (bold_signal.py:97).  ``lbda=None``
    (noise-constrained lambda, bold_signal.py:99-214) estimates the noise level like the reference
    (db3 MAD, utils.py:16-25, on the device) unless ``sigma=`` is given.
    Extra keyword arguments: ``x0`` (power-iteration start), ``sigma``, ``dtype``.  float64 input
    (the reference's type) selects the FP64 build, float32 input or ``dtype=np.float32`` the FP32 one.
    """
    dtype = pick_dtype(y, hrf, dtype=dtype)
    yb, one_d = _as_batch(y, dtype)
    V, T = yb.shape
    hd = to_device(hrf, dtype)
    H = ConvAndLinear(DiscretInteg(), hd, dim_in=T, dim_out=T)
    # bold_signal.py:52 -- 0.9 x power-iteration estimate, start vector from the global RNG
    if x0 is None:
        x0 = np.random.randn(T)
    est = spectral_radius_est(H, (T,), x0=to_device(x0, dtype))
    lipschitz = _lipschitz_cst(est, dtype)

    if lbda is None:
        from .noise import deconv_auto_lbda
        return deconv_auto_lbda(y, yb, one_d, hd, lipschitz, sigma, early_stopping, tol, wind,
                                nb_iter, nb_sub_iter)

    x, z, dz, J, n_iter = deconv_batch(yb, hd, lbda, lipschitz, None, early_stopping, tol, wind,
                                       nb_iter)
    Jn = J / (J[:, :1] + 1.0e-30)                                  # bold_signal.py:97
    if one_d:
        n = int(n_iter[0])
        return (like_input(x[0], y), like_input(z[0], y), like_input(dz[0], y),
                like_input(Jn[0, :n], y), None, None)
    return (like_input(x, y), like_input(z, y), like_input(dz, y), like_input(Jn, y), None, None)


def deconv_lbda_path(y, t_r, hrf, lbdas, nb_iter=200, x0=None, dtype=None, max_problems=1 << 21):
    """Regularisation path: ``deconv`` (fixed lambda, no early stopping) for every lambda of ``lbdas``
    and every voxel of ``y[V, T]`` -- the grid search of examples/icassp_2019/validation.py batched
    over (lambda, voxel) instead of looping (BASELINE.json configs[4]).

    Returns ``(x, z, diff_z, J)`` with a leading lambda axis: ``[n_lbda, V, T]`` and
    ``[n_lbda, V, nb_iter]`` (J normalised by its first entry like ``deconv``).  Every (lambda, voxel)
    pair is an independent problem of ONE persistent launch (``pb_deconv_lbda_path_*``) that reads the
    rows of ``y`` in place -- nothing is replicated per lambda -- and writes straight into the results;
    more than ``max_problems`` pairs are cut into launches of whole lambdas.
    """
    dtype = pick_dtype(y, hrf, dtype=dtype)
    yb, _ = _as_batch(y, dtype)
    yb = yb.contiguous()
    V, T = yb.shape
    hd = to_device(hrf, dtype).reshape(-1).contiguous()
    K = hd.numel()
    lb = upload(np.asarray(lbdas, dtype=np.float64).reshape(-1), dtype, yb.device)
    n_l = lb.numel()
    if x0 is None:
        x0 = np.random.randn(T)
    est = spectral_radius_est(ConvAndLinear(DiscretInteg(), hd, dim_in=T), (T,), x0=to_device(x0, dtype))
    lipschitz = per_voxel(_lipschitz_cst(est, dtype), 1, dtype, yb.device, "lipschitz")[0]
    outs = [torch.empty((n_l, V, T), dtype=dtype, device=yb.device) for _ in range(3)]
    J = torch.empty((n_l, V, nb_iter), dtype=dtype, device=yb.device)
    n_iter = torch.empty(n_l * V, dtype=torch.int32, device=yb.device)
    per = max(1, min(n_l, max_problems // max(V, 1)))
    with torch.cuda.device(yb.device):
        for l0 in range(0, n_l, per):
            l1 = min(l0 + per, n_l)
            rc = _lib.fn("pb_deconv_lbda_path", dtype)(
                ptr(yb), ptr(hd), ptr(lipschitz), ptr(lb[l0:l1]), l1 - l0, int(nb_iter),
                ptr(outs[0][l0:l1]), ptr(outs[1][l0:l1]), ptr(outs[2][l0:l1]), ptr(J[l0:l1]),
                ptr(n_iter[l0 * V:l1 * V]), V, T, K, stream_ptr())
            _lib.check(rc, "pb_deconv_lbda_path")
    J /= (J[:, :, :1] + 1.0e-30)                                   # bold_signal.py:97, in place
    return tuple(like_input(t, y) for t in outs) + (like_input(J, y),)


def bd_alloc(V, T, K, nb_iter, dtype, dev):
    """Output buffers of :func:`bd_batch` (reusable across calls via ``out=``)."""
    return {
        "x": torch.empty((V, T), dtype=dtype, device=dev),
        "z": torch.empty((V, T), dtype=dtype, device=dev),
        "diff_z": torch.empty((V, T), dtype=dtype, device=dev),
        "h": torch.empty((V, K), dtype=dtype, device=dev),
        "theta": torch.empty(V, dtype=dtype, device=dev),
        "J": torch.full((V, nb_iter + 2), float("nan"), dtype=dtype, device=dev),
        "r": torch.full((V, nb_iter + 2), float("nan"), dtype=dtype, device=dev),
        "g": torch.full((V, nb_iter + 2), float("nan"), dtype=dtype, device=dev),
        "n_trace": torch.zeros(V, dtype=torch.int32, device=dev),
    }


def bd_batch(y, t_r, lbda, theta_0, z_0, hrf_dur, bounds, nb_iter, early_stopping, wind, tol,
             out=None):
    """Device entry point of :func:`bd`: tensors in, dict of tensors out (all ``[V, ...]``).

    ``out`` (from :func:`bd_alloc`) lets a caller reuse the output buffers; ``lbda`` / ``theta_0``
    given as device tensors are used in place (no host round trip): the call is then a single
    asynchronous kernel launch.
    """
    V, T = y.shape
    dtype, dev = y.dtype, y.device
    K = hrf_len(t_r, hrf_dur)
    lb, lb_stride = per_voxel(lbda, V, dtype, dev, "lbda")
    th0, th_stride = per_voxel(theta_0, V, dtype, dev, "theta_0")
    lo, hi = bounds[0]
    if out is None:
        out = bd_alloc(V, T, K, nb_iter, dtype, dev)
    if z_0 is not None and (tuple(z_0.shape) != (V, T) or z_0.dtype != dtype or z_0.device != dev):
        raise ValueError("z_0 must be a %s tensor of shape (%d, %d) on %s" % (dtype, V, T, dev))
    with torch.cuda.device(dev):        # grid sizing and the launch stream follow the current device
        rc = _lib.fn("pb_bd", dtype)(
            ptr(y), float(t_r), float(hrf_dur), ptr(lb), lb_stride, ptr(th0), th_stride, ptr(z_0),
            float(lo), float(hi), int(nb_iter), int(bool(early_stopping)), int(wind), float(tol),
            ptr(out["x"]), ptr(out["z"]), ptr(out["diff_z"]), ptr(out["h"]), ptr(out["theta"]),
            ptr(out["J"]), ptr(out["r"]), ptr(out["g"]), ptr(out["n_trace"]), V, T, K, stream_ptr())
    _lib.check(rc, "pb_bd")
    return out


# ---- host-resident batches: chunked, double-buffered streaming ------------------------------
_STREAM_MIN_VOXELS = 8192      # below this a single launch is used
_STREAM_TARGET_CHUNK = 32768   # approximate chunk size, rounded to whole waves of the grid
_PINNED_RESULT_LIMIT = 4 << 30  # results up to this many bytes are returned in pinned host memory
_staging_cache = {}


def _staging(key, shapes, dtype):
    """Pinned staging buffers, cached across calls (cudaHostAlloc is expensive).  The cache is keyed
    by thread and device: one thread per GPU in the same process must not share staging slots."""
    import threading
    key = (threading.get_ident(), torch.cuda.current_device()) + tuple(key)
    bufs = _staging_cache.get(key)
    if bufs is None:
        if len(_staging_cache) > 8:
            _staging_cache.clear()
        bufs = {k: torch.empty(shp, dtype=(torch.int32 if k == "n_trace" else dtype), pin_memory=True)
                for k, shp in shapes.items()}
        _staging_cache[key] = bufs
    return bufs


def _bd_streamed(yh, dtype, t_r, lbda, theta_0, z_0, hrf_dur, bounds, nb_iter, early_stopping, wind,
                 tol):
    """``bd`` for a HOST ``[V, T]`` batch: the batch is cut into chunks of whole grid waves; while
    the kernel solves chunk i, chunk i+1 is uploaded and chunk i-1 is downloaded (two CUDA streams,
    pinned staging), so that the host<->device copies hide behind the solve.  Returns CPU tensors.
    """
    V, T = yh.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    K = hrf_len(t_r, hrf_dur)
    ntr = nb_iter + 2
    # Chunks of whole waves of the persistent grid (no partially filled last wave inside a chunk).  Equal,
    # non-wave-aligned chunks were measured in round 2 and rejected: the kernels of consecutive chunks (two
    # streams) do not overlap at their tails, 4 x 25 000 voxels took 631 ms against 591-599 for 3 x 31 968 + 4 096.
    wave = _lib.lib.pb_bd_wave_voxels(T, K, int(dtype == torch.float64), int(nb_iter))
    chunk = _STREAM_TARGET_CHUNK
    if wave > 0:
        chunk = max(wave, (chunk // wave) * wave)
    chunk = min(chunk, V)
    shapes = {"x": (chunk, T), "z": (chunk, T), "diff_z": (chunk, T), "h": (chunk, K),
              "theta": (chunk,), "J": (chunk, ntr), "r": (chunk, ntr), "g": (chunk, ntr),
              "n_trace": (chunk,)}
    # Results of moderate size are written by the device straight into pinned host tensors (PyTorch
    # caches pinned blocks, so repeated calls do not pay cudaHostAlloc again); larger ones go through
    # two pinned staging slots and a host copy into pageable memory.
    esize = torch.empty((), dtype=dtype).element_size()
    out_bytes = sum(V * int(np.prod(shp[1:], dtype=np.int64)) * esize for shp in shapes.values())
    pin_out = out_bytes <= _PINNED_RESULT_LIMIT

    def alloc_final(pinned):
        return {k: torch.empty((V,) + shp[1:], dtype=(torch.int32 if k == "n_trace" else dtype),
                               pin_memory=pinned)
                for k, shp in shapes.items()}

    try:
        final = alloc_final(pin_out)
    except RuntimeError:            # the host refuses to lock that much memory: pageable + staging
        if not pin_out:
            raise
        pin_out = False
        final = alloc_final(False)

    def host_vec(val):
        if isinstance(val, torch.Tensor):
            val = val.detach().cpu().numpy()
        return np.asarray(val, dtype=np.float64).reshape(-1)

    lb_all, th_all = host_vec(lbda), host_vec(theta_0)
    z0h = None
    if z_0 is not None:
        z0h = (z_0 if isinstance(z_0, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(z_0)))
        z0h = z0h.reshape(V, T)
    slots = []
    for s_idx in range(2):
        slots.append({
            "stream": torch.cuda.Stream(device=dev),
            "event": torch.cuda.Event(),
            "y": torch.empty((chunk, T), dtype=dtype, device=dev),
            "z0": torch.empty((chunk, T), dtype=dtype, device=dev) if z0h is not None else None,
            "out": bd_alloc(chunk, T, K, nb_iter, dtype, dev),
            "stage_in": _staging(("in", s_idx, chunk, T, dtype, z0h is not None),
                                 {"y": (chunk, T), "z0": (chunk, T)} if z0h is not None else {"y": (chunk, T)},
                                 dtype),
            "stage_out": None if pin_out else _staging(("out", s_idx, chunk, T, K, ntr, dtype), shapes, dtype),
            "range": None,
        })

    def drain(slot):
        if slot["range"] is None:
            return
        lo, hi = slot["range"]
        slot["event"].synchronize()
        if not pin_out:
            for k in final:
                final[k][lo:hi].copy_(slot["stage_out"][k][:hi - lo])
        slot["range"] = None

    main = torch.cuda.current_stream()
    n_chunks = (V + chunk - 1) // chunk
    for i in range(n_chunks):
        lo, hi = i * chunk, min((i + 1) * chunk, V)
        n = hi - lo
        slot = slots[i % 2]
        drain(slot)
        src = yh[lo:hi]
        if not (src.is_pinned() and src.dtype == dtype):
            slot["stage_in"]["y"][:n].copy_(src)
            src = slot["stage_in"]["y"][:n]
        z0src = None
        if z0h is not None:
            slot["stage_in"]["z0"][:n].copy_(z0h[lo:hi])
            z0src = slot["stage_in"]["z0"][:n]
        st = slot["stream"]
        st.wait_stream(main)
        with torch.cuda.stream(st):
            slot["y"][:n].copy_(src, non_blocking=True)
            if z0src is not None:
                slot["z0"][:n].copy_(z0src, non_blocking=True)
            lb = lb_all if lb_all.size == 1 else lb_all[lo:hi]
            th = th_all if th_all.size == 1 else th_all[lo:hi]
            out = {k: v[:n] for k, v in slot["out"].items()}
            bd_batch(slot["y"][:n], t_r, lb, th, slot["z0"][:n] if z0src is not None else None,
                     hrf_dur, bounds, nb_iter, early_stopping, wind, tol, out=out)
            for k in final:
                dst = final[k][lo:hi] if pin_out else slot["stage_out"][k][:n]
                dst.copy_(out[k], non_blocking=True)
            slot["event"].record(st)
        slot["range"] = (lo, hi)
    for slot in slots:
        drain(slot)
    main.wait_stream(slots[0]["stream"])
    main.wait_stream(slots[1]["stream"])
    return final


def bd(y, t_r, lbda=1.0, theta_0=None, z_0=None, hrf_dur=20.0,  # noqa
       bounds=None, nb_iter=100, nb_sub_iter=1000, nb_last_iter=10000,
       print_period=50, early_stopping=False, wind=4, tol=1.0e-12, verbose=0, dtype=None):
    """Semi-blind deconvolution with the dilated SPM HRF (pybold/bold_signal.py:281-382).

    Precision follows the input like NumPy would: float64 arrays (what the reference computes in,
    bold_signal.py:288) run the FP64 build -- the parity build, about 2.3x slower -- and float32
    arrays the FP32 build the throughput figures are quoted on (1e-4 of the reference);
    ``dtype=np.float32`` forces the fast build for float64 input.

    Returns ``(x, z, diff_z, h, d)``; ``d`` has the reference's keys ``'J'``, ``'r'``, ``'g'``
    (length ``nb_iter + 2``, shorter after an early stop) and ``'l_alpha'`` (empty list), plus
    ``'theta'`` (final dilation).  Like the reference, ``nb_iter`` is also the inner iteration
    count and ``nb_sub_iter`` / ``nb_last_iter`` are accepted and ignored (bold_signal.py:324,366).
    """
    dtype = pick_dtype(y, dtype=dtype)
    theta_0 = MAX_DELTA if theta_0 is None else theta_0                    # bold_signal.py:291
    th_chk = np.asarray(theta_0.detach().cpu() if isinstance(theta_0, torch.Tensor) else theta_0,
                        dtype=np.float64)
    if np.any(th_chk < MIN_DELTA) or np.any(th_chk > MAX_DELTA):           # hrf_model.py:17-21
        raise ValueError("delta should belong in [{0}, {1}], got delta = {2}".format(
            MIN_DELTA, MAX_DELTA, th_chk))
    if bounds is None:
        bounds = [(MIN_DELTA + 1.0e-1, MAX_DELTA - 1.0e-1)]                # bold_signal.py:303-304
    host_in = not (isinstance(y, torch.Tensor) and y.is_cuda)
    if host_in and np.ndim(y) == 2 and len(y) >= _STREAM_MIN_VOXELS:
        from ._array import require_cuda
        require_cuda()
        yh = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
        out = _bd_streamed(yh.contiguous(), dtype, t_r, lbda, theta_0, z_0, hrf_dur, bounds, nb_iter,
                           early_stopping, wind, tol)
        conv = (lambda t: t) if isinstance(y, torch.Tensor) else (lambda t: t.numpy())
        d = {k: conv(out[k]) for k in ("J", "r", "g", "theta", "n_trace")}
        d["l_alpha"] = []
        return conv(out["x"]), conv(out["z"]), conv(out["diff_z"]), conv(out["h"]), d
    yb, one_d = _as_batch(y, dtype)
    z0 = None
    if z_0 is not None:
        z0 = to_device(z_0, dtype).reshape(yb.shape)
    out = bd_batch(yb, t_r, lbda, theta_0, z0, hrf_dur, bounds, nb_iter, early_stopping, wind, tol)
    if one_d:
        n = int(out["n_trace"][0])
        d = {k: like_input(out[k][0, :n], y) for k in ("J", "r", "g")}
        d["l_alpha"] = []
        d["theta"] = float(out["theta"][0])
        return (like_input(out["x"][0], y), like_input(out["z"][0], y),
                like_input(out["diff_z"][0], y), like_input(out["h"][0], y), d)
    d = {k: like_input(out[k], y) for k in ("J", "r", "g", "theta", "n_trace")}
    d["l_alpha"] = []
    return (like_input(out["x"], y), like_input(out["z"], y), like_input(out["diff_z"], y),
            like_input(out["h"], y), d)


def hrf_estim_batch(z, y, t_r, dur, theta_0=MAX_DELTA, bounds=None):
    """Device entry point: bounded theta minimisation for ``[V, T]`` tensors."""
    V, T = y.shape
    dtype, dev = y.dtype, y.device
    K = hrf_len(t_r, dur)
    if bounds is None:
        bounds = [(MIN_DELTA + 1.0e-1, MAX_DELTA - 1.0e-1)]               # bold_signal.py:229
    th0, th_stride = per_voxel(theta_0, V, dtype, dev, "theta_0")
    theta = torch.empty(V, dtype=dtype, device=dev)
    h = torch.empty((V, K), dtype=dtype, device=dev)
    cost = torch.empty(V, dtype=dtype, device=dev)
    rc = _lib.fn("pb_hrf_estim", dtype)(
        ptr(z), ptr(y), float(t_r), float(dur), ptr(th0), th_stride, float(bounds[0][0]),
        float(bounds[0][1]), ptr(theta), ptr(h), ptr(cost), V, T, K, stream_ptr())
    _lib.check(rc, "pb_hrf_estim")
    return theta, h, cost


def hrf_estim(z, y, t_r, dur, verbose=0):
    """HRF estimation for a known block signal (pybold/bold_signal.py:225-239).

    Returns ``(h, J)``.  The reference's J lists the cost after every L-BFGS-B iterate (its
    ``Tracker`` callback); the device solver is not an iterate-by-iterate port, so J holds the
    cost at the start point and at the minimiser.
    """
    dtype = pick_dtype(z, y)
    yb, one_d = _as_batch(y, dtype)
    zb, _ = _as_batch(z, dtype)
    theta, h, cost = hrf_estim_batch(zb, yb, t_r, dur)
    start_cost = hrf_fit_err(MAX_DELTA - 1.0e-1, z, y, t_r, dur)
    if one_d:
        return like_input(h[0], y), [float(np.asarray(start_cost).reshape(-1)[0]), float(cost[0])]
    return like_input(h, y), [like_input(torch.as_tensor(start_cost), y), like_input(cost, y)]


def hrf_fit_err(theta, z, y, t_r, hrf_dur):
    """0.5 * || y - h(theta) * z ||^2 (pybold/bold_signal.py:217-222), evaluated on the device."""
    from .convolution import spectral_convolve
    from .hrf_model import spm_hrf
    dtype = pick_dtype(z, y)
    yb, one_d = _as_batch(y, dtype)
    zb, _ = _as_batch(z, dtype)
    th = torch.as_tensor(np.asarray(theta, dtype=np.float64).reshape(-1))
    h, _ = spm_hrf(th.to(yb.device), t_r, hrf_dur, False)
    h = h.to(dtype)
    res = yb - spectral_convolve(h if h.shape[0] > 1 else h[0], zb)
    val = 0.5 * torch.sum(res * res, dim=1)
    if one_d:
        return float(val[0])
    return like_input(val, y)
