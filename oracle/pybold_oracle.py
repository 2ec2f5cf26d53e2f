"""CPU oracle for the pyBOLD hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a plain NumPy/SciPy float64 restatement of the algorithms executed by
``pybold.bold_signal.deconv`` / ``bd`` in the reference (hcherkaoui/pybold).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the shipped package ``pybold_b200`` never
does (``tests/test_boundary.py`` greps for that).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the live reference
from ``/root/reference`` (with a stub ``pywt`` and ``np.float``) and stores its
outputs for ``deconv`` (fixed lambda), ``_loops_deconv``, ``bd``, ``spm_hrf``,
``spectral_radius_est``, ``hrf_fit_err`` and the linear operators in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function of
this file against those vectors (bit-exact or <=1e-15 relative for everything
except the L-BFGS-B theta step, which calls the very same SciPy routine).
The single unpinned item is ``mad_daub_noise_est`` (PyWavelets is not installed, so
the reference's sigma cannot be produced here); its header says so.

Every function cites the reference ``file:line`` (relative to the reference root)
it restates.  The recursions are written explicitly (``u``, ``v``, ``w``) instead
of relying on the NumPy aliasing the reference depends on.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.optimize import fmin_l_bfgs_b
from scipy.stats import gamma as _gamma

# step tolerances of the exact theta solver (same constants on the device): a Newton step of size dx
# leaves an error ~dx^2, a bisection step ~dx
THETA_XTOL_NEWTON = 1.0e-7
THETA_XTOL_BISECT = 1.0e-12
MIN_DELTA = 0.5   # pybold/hrf_model.py:8
MAX_DELTA = 2.0   # pybold/hrf_model.py:9


# --------------------------------------------------------------------------------------
# A1 / A2 / A5: linear operators
# --------------------------------------------------------------------------------------
def integ_op(x):
    """``DiscretInteg.op``: running sum (pybold/linear.py:15-28)."""
    return np.cumsum(np.asarray(x, dtype=np.float64))


def integ_adj(x):
    """``DiscretInteg.adj``: out[i] = sum_{j>=i} x[j] (pybold/linear.py:30-43)."""
    x = np.asarray(x, dtype=np.float64)
    return np.cumsum(x[::-1])[::-1]


def conv_causal(k, x):
    """out[i] = sum_j k[j] x[i-j], i < len(x).

    Semantics of ``simple_convolve`` (pybold/convolution.py:135-164), of the Toeplitz
    product ``K.dot(x)`` (convolution.py:105-132, linear.py:91) and -- for every length
    that is not a power of two >= 1024 -- of ``spectral_convolve`` (convolution.py:9-30).
    """
    k = np.asarray(k, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(len(x))
    for j in range(min(len(k), len(x))):
        out[j:] += k[j] * x[:len(x) - j]
    return out


def corr_anticausal(k, x):
    """out[i] = sum_j k[j] x[i+j].

    Semantics of ``simple_retro_convolve`` (pybold/convolution.py:167-196), ``K.T.dot(x)``
    (linear.py:111) and ``spectral_retro_convolve`` (convolution.py:33-54).
    """
    k = np.asarray(k, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(len(x))
    for j in range(min(len(k), len(x))):
        out[:len(x) - j] += k[j] * x[j:]
    return out


def toeplitz_from_kernel(k, dim):
    """Dense square Toeplitz matrix K[i, j] = k[i-j] (pybold/convolution.py:105-132)."""
    k = np.asarray(k, dtype=np.float64)
    K = np.zeros((dim, dim))
    for lag in range(min(len(k), dim)):
        K[np.arange(lag, dim), np.arange(0, dim - lag)] = k[lag]
    return K


class HrfIntegOperator:
    """``ConvAndLinear(DiscretInteg(), hrf, T, T)``: op = K @ cumsum, adj = revcumsum(K.T @ .)

    Follows pybold/linear.py:49-113 with the dense Toeplitz matrix, exactly like the
    reference (this is also what makes the CPU baseline cost representative).
    """

    def __init__(self, hrf, dim):
        self.K = toeplitz_from_kernel(hrf, dim)
        self.K_T = self.K.T

    def op(self, x):
        return self.K.dot(np.cumsum(x))

    def adj(self, x):
        return integ_adj(self.K_T.dot(x))


# --------------------------------------------------------------------------------------
# A9: SPM HRF with dilation
# --------------------------------------------------------------------------------------
def spm_hrf(delta, t_r=1.0, dur=60.0, normalized_hrf=True, dt=0.001, p_delay=6,
            undershoot=16.0, p_disp=1.0, u_disp=1.0, p_u_ratio=0.167, onset=0.0):
    """Fine-grid evaluation then striding, as pybold/hrf_model.py:12-39 does it."""
    delta = float(np.asarray(delta).reshape(-1)[0])
    if delta < MIN_DELTA or delta > MAX_DELTA:           # hrf_model.py:17-21
        raise ValueError("delta should belong in [{0}, {1}], got {2}".format(
            MIN_DELTA, MAX_DELTA, delta))
    n_fine = int(float(dur) / dt)
    t = np.linspace(0, dur, n_fine) - float(onset) / dt  # hrf_model.py:25
    s = delta * t
    hrf = (_gamma.pdf(s, p_delay / p_disp, loc=dt / p_disp)
           - p_u_ratio * _gamma.pdf(s, undershoot / u_disp, loc=dt / u_disp))
    if normalized_hrf:
        hrf = hrf / np.max(hrf + 1.0e-30)                # hrf_model.py:33-34
    stride = int(t_r / dt)
    return hrf[::stride], t[::stride]


def hrf_len(t_r, dur, dt=0.001):
    """K = number of kept samples of the strided fine grid (hrf_model.py:36)."""
    return len(range(0, int(float(dur) / dt), int(t_r / dt)))


def spm_hrf_closed_form(delta, t_r, dur, dt=0.001):
    """Non-normalised taps in closed form at the kept samples only.

    t_m = m * stride * dur / (N - 1), s = delta * t_m - dt,
    h_m = g_6(s) - 0.167 g_16(s), g_a(s) = s^(a-1) e^(-s) / Gamma(a) for s > 0 else 0.
    This is what the device evaluates; tests pin it against :func:`spm_hrf`.
    """
    n_fine = int(float(dur) / dt)
    stride = int(t_r / dt)
    m = np.arange(0, n_fine, stride, dtype=np.float64)
    t = m * (float(dur) / (n_fine - 1))
    s = float(delta) * t - dt
    pos = s > 0
    sp = np.where(pos, s, 1.0)
    g6 = sp ** 5 * np.exp(-sp) / math.factorial(5)
    g16 = sp ** 15 * np.exp(-sp) / math.factorial(15)
    return np.where(pos, g6 - 0.167 * g16, 0.0)


def _gamma_pdf(s, a, loc):
    """scipy.stats.gamma.pdf(s, a, loc=loc) in closed form: x^(a-1) e^(-x) / Gamma(a), x = s - loc."""
    x = np.asarray(s, dtype=np.float64) - loc
    pos = x > 0
    xp = np.where(pos, x, 1.0)
    val = np.exp((a - 1.0) * np.log(xp) - xp - math.lgamma(a))
    at_zero = 0.0 if a > 1.0 else (1.0 if a == 1.0 else np.inf)
    return np.where(pos, val, np.where(x == 0, at_zero, 0.0))


def spm_hrf_general(delta, t_r=1.0, dur=60.0, normalized_hrf=True, dt=0.001, p_delay=6,
                    undershoot=16.0, p_disp=1.0, u_disp=1.0, p_u_ratio=0.167, onset=0.0):
    """pybold/hrf_model.py:12-39 for ANY shape parameters, evaluated only where it is needed (the
    kept samples, plus one pass over the fine grid for the normalisation maximum) -- the statement of
    what ``pb_spm_hrf_ex_*`` computes.  Note the reference's ``t = linspace(...) - onset / dt``
    (hrf_model.py:25): the onset is divided by dt, so ``onset = 0.004`` shifts the HRF by 4 s."""
    n_fine = int(float(dur) / dt)
    stride = int(t_r / dt)
    t_step = float(dur) / (n_fine - 1)

    def at(idx):
        t = idx.astype(np.float64) * t_step - float(onset) / dt
        s = float(delta) * t
        return (_gamma_pdf(s, p_delay / p_disp, dt / p_disp)
                - p_u_ratio * _gamma_pdf(s, undershoot / u_disp, dt / u_disp)), t

    h, t = at(np.arange(0, n_fine, stride))
    if normalized_hrf:
        fine, _ = at(np.arange(n_fine))
        h = h / np.max(fine + 1.0e-30)
    return h, t


# --------------------------------------------------------------------------------------
# A4: power-iteration Lipschitz estimate
# --------------------------------------------------------------------------------------
def spectral_radius_est(H, x0, nb_iter=30, tol=1.0e-6):
    """pybold/utils.py:94-109 with the random start made an explicit argument."""
    x_old = np.array(x0, dtype=np.float64)
    x_new = x_old
    for _ in range(nb_iter):
        x_new = H.adj(H.op(x_old)) / np.linalg.norm(x_old)
        if abs(np.linalg.norm(x_new) - np.linalg.norm(x_old)) < tol:
            break
        x_old = x_new
    return np.linalg.norm(x_new)


def frobenius_lipschitz(hrf, dim):
    """||A^T A||_F with A = K @ tril(1) (pybold/bold_signal.py:249-253), dense."""
    A = toeplitz_from_kernel(hrf, dim).dot(np.tril(np.ones((dim, dim))))
    return np.linalg.norm(A.T.dot(A))


# --------------------------------------------------------------------------------------
# shared prox-gradient pieces
# --------------------------------------------------------------------------------------
def _soft(u, th):
    return np.sign(u) * np.maximum(np.abs(u) - th, 0)


def momentum_weights(nb_iter):
    """beta_k = (t_{k-1} - 1) / t_k with t_{-1} = 1, t_k = (1 + sqrt(1 + 4 t_{k-1}^2)) / 2.

    pybold/bold_signal.py:68-71 (and :119-122, :190-193, :264-275).  beta_0 = 0.
    """
    beta = np.empty(nb_iter)
    t_old = 1.0
    for k in range(nb_iter):
        t = 0.5 * (1.0 + math.sqrt(1.0 + 4.0 * t_old * t_old))
        beta[k] = (t_old - 1.0) / t
        t_old = t
    return beta


# --------------------------------------------------------------------------------------
# A3: deconv, fixed lambda
# --------------------------------------------------------------------------------------
def deconv_fixed_lbda(y, hrf, lbda, x0_power=None, lipschitz=None, early_stopping=True,
                      tol=1.0e-6, wind=6, nb_iter=1000):
    """pybold/bold_signal.py:49-97, written with the recursion that actually executes.

    ``u = w - step (A^T A w - A^T y)``, ``v = soft(u, th)``, ``w <- v + beta_k (v - p)``
    with ``p = 0`` at k = 0 and ``p = u`` afterwards.  ``lipschitz`` is
    the value of ``0.9 * spectral_radius_est`` when supplied, else it is computed from
    ``x0_power`` (the reference draws that start from the global RNG, utils.py:97).

    Returns (x, z, w, J / (J[0] + 1e-30), n_done).
    """
    y = np.asarray(y, dtype=np.float64)
    T = len(y)
    H = HrfIntegOperator(hrf, T)
    Aty = H.adj(y)
    if lipschitz is None:
        lipschitz = 0.9 * spectral_radius_est(H, x0_power)
    step = 1.0 / lipschitz
    th = lbda / lipschitz
    w = np.zeros(T)
    t_old = 1.0
    J = []
    hist = []            # what the reference's list `xx` holds (Q5)
    sub = int(wind / 2)
    x = z = None
    for k in range(nb_iter):
        grad = H.adj(H.op(w)) - Aty
        u = w - step * grad
        v = _soft(u, th)
        t = 0.5 * (1.0 + math.sqrt(1.0 + 4.0 * t_old * t_old))
        p = np.zeros(T) if k == 0 else u
        w = v + (t_old - 1.0) / t * (v - p)
        t_old = t
        z = np.cumsum(w)
        x = conv_causal(hrf, z)
        J.append(0.5 * np.sum(np.square(x - y)) + lbda * np.sum(np.abs(w)))
        # `xx.append(diff_z_old)` binds the array that the next iteration turns into u_{k+1}:
        # at test time the list reads [..., u_{k-1}, u_k, w_k].
        if hist:
            hist[-1] = u
        hist.append(w)
        if len(hist) > wind:
            hist = hist[1:]
        if early_stopping and k > wind:
            old_iter = np.mean(hist[:-sub], axis=0)
            new_iter = np.mean(hist[-sub:], axis=0)
            crit = np.linalg.norm(new_iter - old_iter) / (np.linalg.norm(new_iter) + 1.0e-10)
            if crit < tol:
                break
    J = np.array(J)
    return x, z, w, J / (J[0] + 1.0e-30), len(J)


# --------------------------------------------------------------------------------------
# A6: inner loop of bd
# --------------------------------------------------------------------------------------
def loops_deconv(y, w0, hrf, lbda, nb_iter, early_stopping=False, tol=1.0e-12,
                 return_lipschitz=False):
    """pybold/bold_signal.py:242-278 (``_loops_deconv``), dense Gram matrix like the reference.

    Same executed recursion as :func:`deconv_fixed_lbda`; step = 1 / ||A^T A||_F; early stop
    (Q6) for j > 2 on ``||w_j - u_j|| / (||w_j|| + 1e-10) < tol``.
    """
    y = np.asarray(y, dtype=np.float64)
    T = len(y)
    A = toeplitz_from_kernel(hrf, T).dot(np.tril(np.ones((T, T))))
    AtA = A.T.dot(A)
    Aty = A.T.dot(y)
    lip = np.linalg.norm(AtA)
    step = 1.0 / lip
    th = lbda / lip
    w = np.array(w0, dtype=np.float64)
    t_old = 1.0
    for j in range(nb_iter):
        u = w - step * (AtA.dot(w) - Aty)
        v = _soft(u, th)
        t = 0.5 * (1.0 + math.sqrt(1.0 + 4.0 * t_old * t_old))
        p = np.zeros(T) if j == 0 else u
        w = v + (t_old - 1.0) / t * (v - p)
        if early_stopping and j > 2:
            crit = np.linalg.norm(w - p) / (np.linalg.norm(w) + 1.0e-10)
            if crit < tol:
                break
        t_old = t
    if return_lipschitz:
        return w, lip
    return w


# --------------------------------------------------------------------------------------
# A10: theta step
# --------------------------------------------------------------------------------------
def hrf_fit_err(theta, z, y, t_r, hrf_dur):
    """0.5 || y - h(theta) * z ||^2 (pybold/bold_signal.py:217-222)."""
    h, _ = spm_hrf(theta, t_r, hrf_dur, False)
    return 0.5 * np.sum(np.square(y - conv_causal(h, z)))


def hrf_fit_err_fast(theta, z, y, t_r, hrf_dur):
    """Same cost with the closed-form taps (used by the exact minimiser below)."""
    h = spm_hrf_closed_form(theta, t_r, hrf_dur)
    return 0.5 * np.sum(np.square(y - conv_causal(h, z)))


def theta_step_lbfgsb(theta_prev, z, y, t_r, hrf_dur, bounds):
    """The reference's theta update: SciPy L-BFGS-B with a finite-difference gradient.

    pybold/bold_signal.py:329-333 (same keyword arguments).  Returns a shape-(1,) array,
    as the reference's ``theta`` becomes after the first call.
    """
    theta, _, _ = fmin_l_bfgs_b(func=hrf_fit_err, x0=theta_prev, args=(z, y, t_r, hrf_dur),
                                bounds=bounds, approx_grad=True, maxiter=999, pgtol=1.0e-12)
    return theta


def hrf_estim(z, y, t_r, dur):
    """pybold/bold_signal.py:225-239: theta from x0 = MAX_DELTA (SciPy clips it to the upper
    bound 1.9), ``maxiter=99999``; J lists the cost at every L-BFGS-B iterate (the reference's
    ``Tracker`` callback, utils.py:27-45).  Returns (h, J, theta)."""
    args = (z, y, t_r, dur)
    bounds = [(MIN_DELTA + 1.0e-1, MAX_DELTA - 1.0e-1)]
    J = []
    theta, _, _ = fmin_l_bfgs_b(func=hrf_fit_err, x0=MAX_DELTA, args=args, bounds=bounds,
                                approx_grad=True, callback=lambda th: J.append(hrf_fit_err(th, *args)),
                                maxiter=99999, pgtol=1.0e-12)
    h, _ = spm_hrf(theta, t_r, dur, False)
    return h, J, float(np.asarray(theta).reshape(-1)[0])


def hrf_estim_exact(z, y, t_r, dur):
    """The device algorithm for the same problem: the exact bounded minimiser from the clipped start.
    Returns (h, [cost at the start, cost at the minimiser], theta)."""
    bounds = [(MIN_DELTA + 1.0e-1, MAX_DELTA - 1.0e-1)]
    theta = theta_step_exact(MAX_DELTA, z, y, t_r, dur, bounds)
    h = spm_hrf_closed_form(theta, t_r, dur)
    return h, [hrf_fit_err_fast(bounds[0][1], z, y, t_r, dur), hrf_fit_err_fast(theta, z, y, t_r, dur)], theta


def hrf_taps_and_derivs(theta, t_r, hrf_dur, dt=0.001):
    """Closed-form taps h(theta) and d/dtheta, d2/dtheta2.

    g_a(s) = s^(a-1) e^(-s) / Gamma(a);  g_a' = g_a ((a-1)/s - 1);
    g_a'' = g_a (((a-1)/s - 1)^2 - (a-1)/s^2);  s = theta t_m - dt  =>  d/dtheta = t_m d/ds.
    """
    n_fine = int(float(hrf_dur) / dt)
    stride = int(t_r / dt)
    m = np.arange(0, n_fine, stride, dtype=np.float64)
    t = m * (float(hrf_dur) / (n_fine - 1))
    s = float(theta) * t - dt
    pos = s > 0
    sp = np.where(pos, s, 1.0)
    e = np.exp(-sp)
    out = []
    g = {6: sp ** 5 * e / math.factorial(5), 16: sp ** 15 * e / math.factorial(15)}
    h = g[6] - 0.167 * g[16]
    d1 = {a: g[a] * ((a - 1) / sp - 1.0) for a in (6, 16)}
    d2 = {a: g[a] * (((a - 1) / sp - 1.0) ** 2 - (a - 1) / sp ** 2) for a in (6, 16)}
    h1 = (d1[6] - 0.167 * d1[16]) * t
    h2 = (d2[6] - 0.167 * d2[16]) * t * t
    for arr in (h, h1, h2):
        out.append(np.where(pos, arr, 0.0))
    return out


def theta_step_exact(theta_prev, z, y, t_r, hrf_dur, bounds, max_iter=100):
    """Bounded 1-D local minimiser of :func:`hrf_fit_err` by root finding on f'(theta).

    This is the *device algorithm* restated on the CPU (the reference's L-BFGS-B answer is
    only within ~1e-8 of this point).  From ``clip(theta_prev)`` walk in
    the descent direction (first step = Newton step when f'' > 0, doubling afterwards) until
    f' changes sign or the bound is reached; then bracketed Newton/bisection on f'.
    With z == 0 the cost is flat (f' == 0) and theta stays at the clipped start, which is
    also what L-BFGS-B returns.
    """
    lo, hi = bounds[0]
    theta = min(max(float(np.asarray(theta_prev).reshape(-1)[0]), lo), hi)
    z = np.asarray(z, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)

    def gc(th):
        h, h1, h2 = hrf_taps_and_derivs(th, t_r, hrf_dur)
        r = conv_causal(h, z) - y
        a1 = conv_causal(h1, z)
        a2 = conv_causal(h2, z)
        return r.dot(a1), a1.dot(a1) + r.dot(a2)

    return bracketed_newton(gc, theta, lo, hi, max_iter)


def bracketed_newton(gc, theta, lo, hi, max_iter=100):
    """Shared by :func:`theta_step_exact`; mirrors ``pb_theta_solve`` in csrc/pb_theta.cuh."""
    g, c = gc(theta)
    if g == 0.0 or not np.isfinite(g):
        return theta
    direction = -1.0 if g > 0 else 1.0
    bound = hi if direction > 0 else lo
    if theta == bound:
        return theta
    # ---- phase 1: find a sign change of f' along the descent direction ----
    a, ga, ca = theta, g, c
    step = abs(ga / ca) if ca > 0 else 0.125 * (hi - lo)
    step = min(max(step, 1.0e-6), 0.25 * (hi - lo))
    b = gb = cb = None
    for _ in range(max_iter):
        cand = a + direction * step
        cand = min(cand, hi) if direction > 0 else max(cand, lo)
        g_c, c_c = gc(cand)
        if g_c == 0.0:
            return cand
        if (g_c > 0) != (ga > 0):
            b, gb, cb = cand, g_c, c_c
            break
        a, ga, ca = cand, g_c, c_c
        if cand == bound:
            return bound
        step *= 2.0
    if b is None:
        return a
    # ---- phase 2: bracketed Newton (falls back to bisection) on f' ----
    if ga < 0:
        xl, xh = a, b
    else:
        xl, xh = b, a
    if abs(ga) < abs(gb):
        x, gx, cx = a, ga, ca
    else:
        x, gx, cx = b, gb, cb
    dx_old = abs(xh - xl)
    dx = dx_old
    for _ in range(max_iter):
        newton_ok = cx > 0 and ((x - xh) * cx - gx) * ((x - xl) * cx - gx) < 0 \
            and abs(2.0 * gx) <= abs(dx_old * cx)
        dx_old = dx
        if newton_ok:
            dx = gx / cx
            x_new = x - dx
        else:
            dx = 0.5 * (xh - xl)
            x_new = xl + dx
        if x_new == x:
            return x
        x = x_new
        if abs(dx) <= (THETA_XTOL_NEWTON if newton_ok else THETA_XTOL_BISECT) * max(1.0, abs(x)):
            return x
        gx, cx = gc(x)
        if gx == 0.0:
            return x
        if gx < 0:
            xl = x
        else:
            xh = x
    return x


# --------------------------------------------------------------------------------------
# A7: bd
# --------------------------------------------------------------------------------------
def bd(y, t_r, lbda=1.0, theta_0=None, z_0=None, hrf_dur=20.0, bounds=None, nb_iter=100,
       early_stopping=False, wind=4, tol=1.0e-12, theta_solver="lbfgsb", trace=None):
    """pybold/bold_signal.py:281-382.

    ``theta_solver``: ``"lbfgsb"`` follows the reference call for call (SciPy);
    ``"exact"`` uses :func:`theta_step_exact` (the algorithm the device runs).
    ``trace`` (optional dict) receives the per-outer-iteration theta, Lipschitz constant
    and warm-start iterate so that stage-wise parity can be gated.
    """
    y = np.asarray(y, dtype=np.float64)                       # :288
    T = len(y)
    theta = MAX_DELTA if theta_0 is None else theta_0          # :291
    h, _ = spm_hrf(theta, t_r, hrf_dur, False)                 # :292
    if z_0 is None:                                            # :294-301
        w = np.zeros(T)
        z = np.zeros(T)
        x = np.zeros(T)
    else:
        z_0 = np.asarray(z_0, dtype=np.float64)
        w = np.append(0, z_0[1:] - z_0[:-1])
        z = z_0
        x = conv_causal(h, z)
    if bounds is None:
        bounds = [(MIN_DELTA + 1.0e-1, MAX_DELTA - 1.0e-1)]    # :303-304
    r_0 = np.sum(np.square(x - y))
    g_0 = np.sum(np.abs(w))
    j_0 = r_0 + lbda * g_0
    d = {'r': [1.0], 'g': [g_0], 'J': [1.0], 'l_alpha': []}
    if trace is not None:
        trace.update(theta=[], lipschitz=[], w_in=[], w_out=[], h=[])
    step_fn = theta_step_lbfgsb if theta_solver == "lbfgsb" else theta_step_exact
    sub = int(wind / 2)
    for idx in range(nb_iter):                                  # :320
        if trace is not None:
            trace['w_in'].append(w.copy())
            trace['h'].append(np.array(h))
        w, lip = loops_deconv(y, w, h, lbda, nb_iter, early_stopping, tol,
                              return_lipschitz=True)            # :324 (nb_iter forwarded, Q3)
        z = np.cumsum(w)
        theta = step_fn(theta, z, y, t_r, hrf_dur, bounds)      # :329-333
        h, _ = spm_hrf(theta, t_r, hrf_dur, False)
        x = conv_causal(h, z)
        r = np.sum(np.square(x - y))
        g = np.sum(np.abs(w))
        d['J'].append((r + lbda * g) / j_0 + 1.0e-30)
        d['r'].append(r / r_0 + 1.0e-30)
        d['g'].append(g)
        if trace is not None:
            trace['theta'].append(float(np.asarray(theta).reshape(-1)[0]))
            trace['lipschitz'].append(lip)
            trace['w_out'].append(w.copy())
        if early_stopping and idx > wind:                       # :350-362 (Q7, signed)
            old_j = np.mean(d['J'][:-sub])
            new_j = np.mean(d['J'][-sub:])
            if (new_j - old_j) / new_j < tol:
                break
    if trace is not None:
        trace['w_in'].append(w.copy())
        trace['h'].append(np.array(h))
    w, lip = loops_deconv(y, w, h, lbda, nb_iter, early_stopping, tol,
                          return_lipschitz=True)                # :366
    z = np.cumsum(w)
    x = conv_causal(h, z)
    r = np.sum(np.square(x - y))
    g = np.sum(np.abs(w))
    d['J'].append((r + lbda * g) / j_0)
    d['r'].append(r / r_0)
    d['g'].append(g)
    if trace is not None:
        trace['lipschitz'].append(lip)
        trace['w_out'].append(w.copy())
    for key in ('J', 'r', 'g'):
        d[key] = np.array(d[key])
    return x, z, w, h, d


# --------------------------------------------------------------------------------------
# A8: deconv with the noise-constrained lambda (lbda=None branch)
# --------------------------------------------------------------------------------------
_DB3_DEC_HI = np.array([-0.3326705529509569, 0.8068915093133388, -0.4598775021193313,
                        -0.13501102001039084, 0.08544127388224149, 0.035226291882100656])


_DB1_DEC_HI = np.array([-0.7071067811865476, 0.7071067811865476])


def dwt_detail_level1(x, dec_hi):
    """Level-1 detail coefficients with PyWavelets' default ``symmetric`` (half-sample) extension:
    ``cD[o] = sum_j dec_hi[j] x_ext[2 o + 1 - j]``, ``len = floor((T + F - 1) / 2)``,
    ``x_ext[-1 - k] = x[k]``, ``x_ext[T + k] = x[T - 1 - k]``.

    PARTLY PINNED: PyWavelets is absent here (pybold/utils.py:22 calls ``pywt.wavedec``).  The
    convention (phase ``2 o + 1``, output length, mirror rule) is checked in
    tests/test_oracle_golden.py against the Haar examples of PyWavelets' documentation
    (``pywt.dwt([1, 2, 3, 4, 5, 6], 'db1')`` and the odd-length ``[1, 2, 3]`` case, quoted from the
    documentation, not generated here); the db3 coefficients are pinned by the filter's defining
    properties (tests/test_boundary.py).  No PyWavelets db3 output vector exists in this repo.
    """
    x = np.asarray(x, dtype=np.float64)
    dec_hi = np.asarray(dec_hi, dtype=np.float64)
    T = len(x)
    F = len(dec_hi)
    ext = np.concatenate([x[:F - 1][::-1], x, x[-(F - 1):][::-1]]) if F > 1 else x
    n_out = (T + F - 1) // 2
    out = np.empty(n_out)
    for o in range(n_out):
        idx = 2 * o + 1 + (F - 1)     # position in ext of x_ext[2o+1]
        out[o] = sum(dec_hi[j] * ext[idx - j] for j in range(F))
    return out


def db3_detail_level1(x):
    """``pywt.wavedec(x, 'db3', level=1)[1]`` restated (see :func:`dwt_detail_level1`)."""
    return dwt_detail_level1(x, _DB3_DEC_HI)


def mad(x, c=0.6744):
    """pybold/utils.py:10-13."""
    return np.median(np.abs(x - np.median(x))) / c


def mad_daub_noise_est(x, c=0.6744):
    """pybold/utils.py:16-25.  Series shorter than 10 scans have ``dwt_max_level == 0``: PyWavelets
    < 1.0 raised ValueError for ``level=1`` there and the reference falls back to ``level=0``, whose
    only coefficient array is the series itself (utils.py:23-24)."""
    x = np.asarray(x, dtype=np.float64)
    cD = x if len(x) < 10 else db3_detail_level1(x)
    return mad(cD, c)


def deconv_auto_lbda(y, hrf, sigma, x0_power=None, lipschitz=None, early_stopping=True,
                     tol=1.0e-6, wind=6, nb_iter=1000, nb_sub_iter=1000):
    """pybold/bold_signal.py:99-214 with the noise level ``sigma`` as an explicit input.

    Inner loop = the same executed recursion, ``t`` reset for every outer iteration, ``w``
    carried over; ``alpha += 1e-4 (||x - y||^2 - T sigma^2)``, ``lbda = 1 / (2 alpha)``;
    outer early stop on the window of alphas (Q11).  Returns (x, z, w, J, R, G).
    """
    y = np.asarray(y, dtype=np.float64)
    T = len(y)
    H = HrfIntegOperator(hrf, T)
    Aty = H.adj(y)
    if lipschitz is None:
        lipschitz = 0.9 * spectral_radius_est(H, x0_power)
    step = 1.0 / lipschitz
    sub = int(wind / 2)
    state = {'w': np.zeros(T), 'p': np.zeros(T)}

    def inner(lbda):
        th = lbda / lipschitz
        w = state['w']
        p_prev = state['p']          # only read at j == 0 (beta_0 = 0 makes it irrelevant)
        t_old = 1.0
        hist = []
        for j in range(nb_sub_iter):
            u = w - step * (H.adj(H.op(w)) - Aty)
            v = _soft(u, th)
            t = 0.5 * (1.0 + math.sqrt(1.0 + 4.0 * t_old * t_old))
            p = p_prev if j == 0 else u
            w = v + (t_old - 1.0) / t * (v - p)
            t_old = t
            if hist:
                hist[-1] = u
            hist.append(w)
            if len(hist) > wind:
                hist = hist[1:]
            if early_stopping and j > wind:
                old_iter = np.mean(hist[:-sub], axis=0)
                new_iter = np.mean(hist[-sub:], axis=0)
                crit = (np.linalg.norm(new_iter - old_iter)
                        / (np.linalg.norm(new_iter) + 1.0e-10))
                if crit < tol:
                    break
        state['w'] = w
        state['p'] = w

    l_alpha, J, R, G = [], [], [], []
    alpha = 1.0
    lbda = 1.0 / (2.0 * alpha)
    mu = 1.0e-4
    for i in range(nb_iter):
        inner(lbda)
        w = state['w']
        z = np.cumsum(w)
        x = conv_causal(hrf, z)
        grad = np.sum(np.square(x - y)) - T * sigma ** 2
        alpha += mu * grad
        lbda = 1.0 / (2.0 * alpha)
        l_alpha.append(alpha)
        if len(l_alpha) > wind:
            l_alpha = l_alpha[1:]
        r = np.sum(np.square(x - y))
        g = np.sum(np.abs(w))
        R.append(r)
        G.append(g)
        J.append(0.5 * r + lbda * g)
        if early_stopping and i > wind:
            old_iter = np.mean(l_alpha[:-sub])
            new_iter = np.mean(l_alpha[-sub:])
            if abs(new_iter - old_iter) / abs(new_iter) < tol:
                break
    inner(lbda)
    w = state['w']
    z = np.cumsum(w)
    x = conv_causal(hrf, z)
    return x, z, w, J, R, G, lbda
