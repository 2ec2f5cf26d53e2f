// Device-side building blocks shared by every solver kernel of pybold_b200 (sm_100a).
//
//  * warp reductions / scans,
//  * closed-form SPM HRF taps and their theta-derivatives       (pybold/hrf_model.py:12-39),
//  * the Frobenius-norm Lipschitz constant without the Gram matrix (pybold/bold_signal.py:249-253),
//  * the bounded theta step on 0.5||y - h(theta)*z||^2            (pybold/bold_signal.py:217-222, :329-334).
//
// Everything in this file runs in double whatever the signal precision is: these are
// per-outer-iteration scalar/K-sized computations (<3 % of the work) that every inner
// iteration depends on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PB_FULL 0xffffffffu
// Stopping rule of the theta solver, on the step dx just taken (relative to max(1, |theta|)):
//   after a Newton step the remaining error is ~dx^2 (quadratic convergence), so |dx| <= 1e-7 leaves
//   ~1e-14; after a bisection step it is ~|dx|, so that needs |dx| <= 1e-12.
// The evaluation noise of f' (double rounding of M h - b) puts a floor of ~1e-14 on dx anyway, and
// the reference's own L-BFGS-B answer is ~1e-8 away from the minimiser.
#define PB_THETA_XTOL_NEWTON 1.0e-7
#define PB_THETA_XTOL_BISECT 1.0e-12

namespace pb {

// ------------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PB_FULL, v, o);
    return v;  // xor butterfly: bit-identical on every lane
}

// exclusive prefix (over lanes) of one value per lane
template <typename T>
__device__ __forceinline__ T warp_excl_scan_up(T v, int lane) {
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(PB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    T ex = __shfl_up_sync(PB_FULL, inc, 1);
    return lane == 0 ? T(0) : ex;
}

// exclusive suffix (sum over higher lanes) of one value per lane
template <typename T>
__device__ __forceinline__ T warp_excl_scan_down(T v, int lane) {
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_down_sync(PB_FULL, inc, d);
        if (lane + d < 32) inc += t;
    }
    T ex = __shfl_down_sync(PB_FULL, inc, 1);
    return lane == 31 ? T(0) : ex;
}

// ------------------------------------------------------------------------------------------------
// SPM HRF with dilation theta, closed form at the kept samples
//   t_m = (m * stride) * dur / (N - 1),  s = theta * t_m - dt,  dt = 1e-3
//   h_m = g_6(s) - 0.167 g_16(s),  g_a(s) = s^(a-1) e^(-s) / Gamma(a)  (s > 0, else 0)
//   d g_a / ds = g_(a-1) - g_a  =>  dh/dtheta = t_m (...),  d2h/dtheta2 = t_m^2 (...)
// ------------------------------------------------------------------------------------------------
struct HrfGrid {
    double t_step;  // dur / (N - 1), N = int(dur / dt): spacing of np.linspace(0, dur, N)
    int stride;     // int(t_r / dt)
    int K;          // number of kept samples
    __device__ __forceinline__ double t(int m) const { return (double)(m * stride) * t_step; }
};

__device__ __forceinline__ void hrf_tap(double theta, double t_m, double &h, double &h1, double &h2) {
    const double s = theta * t_m - 1.0e-3;
    if (!(s > 0.0)) {
        h = h1 = h2 = (s == s) ? 0.0 : s;  // NaN theta propagates
        return;
    }
    const double e = exp(-s);
    const double s2 = s * s, s3 = s2 * s, s4 = s2 * s2, s5 = s4 * s;
    const double s10 = s5 * s5;
    const double g4 = e * s3 * (1.0 / 6.0);
    const double g5 = e * s4 * (1.0 / 24.0);
    const double g6 = e * s5 * (1.0 / 120.0);
    const double g14 = e * (s10 * s3) * (1.0 / 6227020800.0);      // 13!
    const double g15 = e * (s10 * s4) * (1.0 / 87178291200.0);     // 14!
    const double g16 = e * (s10 * s5) * (1.0 / 1307674368000.0);   // 15!
    h = g6 - 0.167 * g16;
    h1 = t_m * ((g5 - g6) - 0.167 * (g15 - g16));
    h2 = t_m * t_m * ((g4 - 2.0 * g5 + g6) - 0.167 * (g14 - 2.0 * g15 + g16));
}

// value only (used for the normalisation maximum over the fine grid)
__device__ __forceinline__ double hrf_value(double theta, double t) {
    const double s = theta * t - 1.0e-3;
    if (!(s > 0.0)) return 0.0;
    const double s5 = s * s * s * s * s;
    return exp(-s) * s5 * (1.0 / 120.0 - 0.167 * (s5 * s5) * (1.0 / 1307674368000.0));
}

// ------------------------------------------------------------------------------------------------
// Per-warp scratch (shared memory, doubles) used by the Lipschitz and theta phases.
// KS = row stride of M, odd so that "lane a reads M[a][b]" is bank-conflict free.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int pb_gram_stride(int kmax) { return kmax | 1; }
__host__ __device__ constexpr int pb_scratch_doubles(int kmax) {
    return 9 * kmax + kmax * pb_gram_stride(kmax);
}

struct ThetaScratch {
    double *hs, *h1s, *h2s;  // taps and derivatives at the current theta
    double *cs, *Ss;         // cumsum(h), cumsum(cumsum(h))
    double *b, *Rz, *zend;   // Z^T y, autocorrelation of z, last K samples of z (reversed)
    double *red;             // spare K doubles
    double *M;               // Gram matrix Z^T Z, [K][KS]
    int KS;
    __device__ __forceinline__ void bind(double *base, int kmax) {
        hs = base;
        h1s = hs + kmax;
        h2s = h1s + kmax;
        cs = h2s + kmax;
        Ss = cs + kmax;
        b = Ss + kmax;
        Rz = b + kmax;
        zend = Rz + kmax;
        red = zend + kmax;
        M = red + kmax;
        KS = pb_gram_stride(kmax);
    }
};

// taps at theta into scratch (all K), warp cooperative
__device__ __forceinline__ void hrf_eval_warp(double theta, const HrfGrid &grid, ThetaScratch &sc,
                                              int lane) {
    for (int a = lane; a < grid.K; a += 32) {
        double h, h1, h2;
        hrf_tap(theta, grid.t(a), h, h1, h2);
        sc.hs[a] = h;
        sc.h1s[a] = h1;
        sc.h2s[a] = h2;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// ||A^T A||_F with A = Toeplitz(h) . tril(1), in O(K^2 + T) instead of the reference's two T^3
// dgemm (pybold/bold_signal.py:249-253).
//
// A is lower-triangular Toeplitz with taps c~[m] = c[min(m, K-1)], c = cumsum(h), C = c[K-1].
// G = A^T A has G[i, i+d] = R_d(T-1-i-d), R_d(n) = sum_{m<=n} c~[m] c~[m+d], hence
//   ||G||_F^2 = sum_d m_d sum_{n=0}^{T-1-d} R_d(n)^2,  m_0 = 1, m_d = 2.
// With E = K-1:  for d >= E, R_d(n) = C * S~(n) (independent of d, S~ = prefix sums of c~);
//                for d <  E, R_d(n) is affine in n once n >= E-1 (both taps are C): closed form.
// Verified against the dense formula to 6e-15 (tests/test_gpu_ops.py::test_frobenius_formula_edge_shapes).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double frob_lipschitz_warp(ThetaScratch &sc, int K, int T, int lane) {
    if (lane == 0) {
        double c = 0.0, S = 0.0;
        for (int m = 0; m < K; ++m) {
            c += sc.hs[m];
            sc.cs[m] = c;
            S += c;
            sc.Ss[m] = S;
        }
    }
    __syncwarp();
    const int E = K - 1;
    const double C = sc.cs[K - 1];
    double total = 0.0;
    for (int d = lane; d < E && d < T; d += 32) {
        double acc = 0.0, ss = 0.0;
        const int nmax = min(E - 1, T - 1 - d);
        for (int m = 0; m <= nmax; ++m) {
            const int md = m + d;
            acc = fma(sc.cs[m], sc.cs[md < K - 1 ? md : K - 1], acc);
            ss = fma(acc, acc, ss);
        }
        const int Q = T - d - E;
        if (Q > 0) {
            const double a = acc, e = C * C, q = (double)Q;
            ss += q * a * a + a * e * q * (q + 1.0) + e * e * q * (q + 1.0) * (2.0 * q + 1.0) / 6.0;
        }
        total += (d == 0 ? 1.0 : 2.0) * ss;
    }
    const double SE = E > 0 ? sc.Ss[E - 1] : 0.0;
    const double dup = (E == 0) ? 1.0 : 0.0;
    for (int n = lane; n < T - E; n += 32) {
        const double St = n < E ? sc.Ss[n] : SE + C * (double)(n - E + 1);
        const double wgt = 2.0 * (double)(T - n - E) - dup;
        const double v = C * St;
        total = fma(wgt * v, v, total);
    }
    total = warp_sum(total);
    __syncwarp();
    return sqrt(total);
}

// ------------------------------------------------------------------------------------------------
// theta step.  f(theta) = 0.5||y - Z h(theta)||^2 with Z the T x K Toeplitz matrix of z:
//   f' = h'^T (M h - b),  f'' = h'^T M h' + h''^T (M h - b),  M = Z^T Z,  b = Z^T y.
// M[a][b] = Rz[|a-b|] - sum_{s=1}^{min(a,b)} zend[a-s] zend[b-s]   (edge-corrected autocorrelation)
// so the T-sized work is done once per outer iteration and every evaluation costs O(K^2).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gram_build_warp(ThetaScratch &sc, int K, int lane) {
    for (int d = lane; d < K; d += 32) {
        double acc = sc.Rz[d];
        sc.M[d] = acc;
        sc.M[d * sc.KS] = acc;
        for (int n = 1; n + d < K; ++n) {
            acc = fma(-sc.zend[n - 1], sc.zend[n - 1 + d], acc);
            sc.M[n * sc.KS + n + d] = acc;
            sc.M[(n + d) * sc.KS + n] = acc;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void theta_eval_warp(double theta, const HrfGrid &grid, ThetaScratch &sc,
                                                int lane, double &g, double &c) {
    hrf_eval_warp(theta, grid, sc, lane);
    const int K = grid.K;
    double pg = 0.0, pc = 0.0;
    for (int a = lane; a < K; a += 32) {
        double q = -sc.b[a], q1 = 0.0;
        const double *row = sc.M + a * sc.KS;
        for (int bb = 0; bb < K; ++bb) {
            const double m = row[bb];
            q = fma(m, sc.hs[bb], q);
            q1 = fma(m, sc.h1s[bb], q1);
        }
        pg = fma(sc.h1s[a], q, pg);
        pc = fma(sc.h1s[a], q1, pc);
        pc = fma(sc.h2s[a], q, pc);
    }
    g = warp_sum(pg);
    c = warp_sum(pc);
    __syncwarp();
}

// Bracketed Newton on f' from clip(theta_prev): walk in the descent direction until f' changes
// sign (or the bound is hit), then safeguarded Newton / bisection.  Mirrors
// the `bracketed_newton` restatement kept with the CPU checker under oracle/ statement for statement.  Control flow is warp
// uniform because g and c come out of xor-butterfly reductions.
__device__ __forceinline__ double theta_solve_warp(double theta_prev, double lo, double hi,
                                                   const HrfGrid &grid, ThetaScratch &sc, int lane,
                                                   int *n_eval) {
    const int max_iter = 100;
    double theta = fmin(fmax(theta_prev, lo), hi);
    double g, c;
    int evals = 1;
    theta_eval_warp(theta, grid, sc, lane, g, c);
    double result = theta;
    bool done = false;
    if (g == 0.0 || !isfinite(g)) done = true;
    const double direction = g > 0.0 ? -1.0 : 1.0;
    const double bound = direction > 0.0 ? hi : lo;
    if (!done && theta == bound) done = true;
    if (!done) {
        double a = theta, ga = g, ca = c;
        double step = ca > 0.0 ? fabs(ga / ca) : 0.125 * (hi - lo);
        step = fmin(fmax(step, 1.0e-6), 0.25 * (hi - lo));
        double b = 0.0, gb = 0.0, cb = 0.0;
        bool have_b = false;
        for (int it = 0; it < max_iter && !done; ++it) {
            double cand = a + direction * step;
            cand = direction > 0.0 ? fmin(cand, hi) : fmax(cand, lo);
            double g_c, c_c;
            theta_eval_warp(cand, grid, sc, lane, g_c, c_c);
            ++evals;
            if (g_c == 0.0) {
                result = cand;
                done = true;
                break;
            }
            if ((g_c > 0.0) != (ga > 0.0)) {
                b = cand;
                gb = g_c;
                cb = c_c;
                have_b = true;
                break;
            }
            a = cand;
            ga = g_c;
            ca = c_c;
            if (cand == bound) {
                result = bound;
                done = true;
                break;
            }
            step *= 2.0;
        }
        if (!done && !have_b) {
            result = a;
            done = true;
        }
        if (!done) {
            double xl, xh;
            if (ga < 0.0) {
                xl = a;
                xh = b;
            } else {
                xl = b;
                xh = a;
            }
            double x, gx, cx;
            if (fabs(ga) < fabs(gb)) {
                x = a;
                gx = ga;
                cx = ca;
            } else {
                x = b;
                gx = gb;
                cx = cb;
            }
            double dx_old = fabs(xh - xl);
            double dx = dx_old;
            result = x;
            for (int it = 0; it < max_iter; ++it) {
                const bool newton_ok = cx > 0.0 &&
                                       ((x - xh) * cx - gx) * ((x - xl) * cx - gx) < 0.0 &&
                                       fabs(2.0 * gx) <= fabs(dx_old * cx);
                dx_old = dx;
                double x_new;
                if (newton_ok) {
                    dx = gx / cx;
                    x_new = x - dx;
                } else {
                    dx = 0.5 * (xh - xl);
                    x_new = xl + dx;
                }
                if (x_new == x) break;
                x = x_new;
                result = x;
                if (fabs(dx) <= (newton_ok ? PB_THETA_XTOL_NEWTON : PB_THETA_XTOL_BISECT) * fmax(1.0, fabs(x))) break;
                theta_eval_warp(x, grid, sc, lane, gx, cx);
                ++evals;
                if (gx == 0.0) break;
                if (gx < 0.0) xl = x; else xh = x;
            }
        }
    }
    if (n_eval) *n_eval = evals;
    return result;
}

// momentum weights beta_k = (t_{k-1} - 1) / t_k, t_k = (1 + sqrt(1 + 4 t_{k-1}^2)) / 2
// (pybold/bold_signal.py:68-71, :264-275).  Data independent: one thread fills a table per CTA.
template <typename real>
__device__ __forceinline__ void fill_momentum_table(real *beta, int nb_iter) {
    if (threadIdx.x == 0) {
        double t_old = 1.0;
        for (int k = 0; k < nb_iter; ++k) {
            const double t = 0.5 * (1.0 + sqrt(1.0 + 4.0 * t_old * t_old));
            beta[k] = (real)((t_old - 1.0) / t);
            t_old = t;
        }
    }
    __syncthreads();
}

}  // namespace pb
