// Group-parallel (lock-step) versions of the double-precision phases of pb_device.cuh, for the
// kernels that put 32/G voxels in one warp (pb_fastg.cuh): every group of G lanes works on its own
// voxel's scratch at the same time, so the latency of the Lipschitz constant, the Gram matrix and
// the theta solve is paid once per warp instead of once per voxel.
//
// The theta solver is the same bracketed Newton as theta_solve_warp (pb_device.cuh) unrolled into a
// state machine: every pass evaluates f'(theta), f''(theta) for all groups (a finished group
// re-evaluates at its answer), then each group advances its own state; no shuffle or barrier sits
// inside the divergent part.  The sequence of evaluation points of a voxel is identical to the
// sequential solver's.
#pragma once
#include "pb_device.cuh"

namespace pb {

template <int G>
__device__ __forceinline__ double group_sum_f64(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(PB_FULL, v, o, G);
    return v;
}

template <int G>
__device__ __forceinline__ void hrf_eval_group(double theta, const HrfGrid &grid, ThetaScratch &sc,
                                               int q) {
    for (int a = q; a < grid.K; a += G) {
        double h, h1, h2;
        hrf_tap(theta, grid.t(a), h, h1, h2);
        sc.hs[a] = h;
        sc.h1s[a] = h1;
        sc.h2s[a] = h2;
    }
    __syncwarp();
}

// same formula as frob_lipschitz_warp (pb_device.cuh), lanes of the group stride over the lags
template <int G>
__device__ __forceinline__ double frob_lipschitz_group(ThetaScratch &sc, int K, int T, int q) {
    if (q == 0) {
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
    for (int d = q; d < E && d < T; d += G) {
        double acc = 0.0, ss = 0.0;
        const int nmax = min(E - 1, T - 1 - d);
        for (int m = 0; m <= nmax; ++m) {
            const int md = m + d;
            acc = fma(sc.cs[m], sc.cs[md < K - 1 ? md : K - 1], acc);
            ss = fma(acc, acc, ss);
        }
        const int Q = T - d - E;
        if (Q > 0) {
            const double a = acc, e = C * C, qq = (double)Q;
            ss += qq * a * a + a * e * qq * (qq + 1.0) + e * e * qq * (qq + 1.0) * (2.0 * qq + 1.0) / 6.0;
        }
        total += (d == 0 ? 1.0 : 2.0) * ss;
    }
    const double SE = E > 0 ? sc.Ss[E - 1] : 0.0;
    const double dup = (E == 0) ? 1.0 : 0.0;
    for (int n = q; n < T - E; n += G) {
        const double St = n < E ? sc.Ss[n] : SE + C * (double)(n - E + 1);
        const double wgt = 2.0 * (double)(T - n - E) - dup;
        const double v = C * St;
        total = fma(wgt * v, v, total);
    }
    total = group_sum_f64<G>(total);
    __syncwarp();
    return sqrt(total);
}

template <int G>
__device__ __forceinline__ void gram_build_group(ThetaScratch &sc, int K, int q) {
    for (int d = q; d < K; d += G) {
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

template <int G>
__device__ __forceinline__ void theta_eval_group(double theta, const HrfGrid &grid, ThetaScratch &sc,
                                                 int q, double &g, double &c) {
    hrf_eval_group<G>(theta, grid, sc, q);
    const int K = grid.K;
    double pg = 0.0, pc = 0.0;
    for (int a = q; a < K; a += G) {
        double qa = -sc.b[a], q1 = 0.0;
        const double *row = sc.M + a * sc.KS;
        for (int bb = 0; bb < K; ++bb) {
            const double m = row[bb];
            qa = fma(m, sc.hs[bb], qa);
            q1 = fma(m, sc.h1s[bb], q1);
        }
        pg = fma(sc.h1s[a], qa, pg);
        pc = fma(sc.h1s[a], q1, pc);
        pc = fma(sc.h2s[a], qa, pc);
    }
    g = group_sum_f64<G>(pg);
    c = group_sum_f64<G>(pc);
    __syncwarp();
}

template <int G>
__device__ __forceinline__ double theta_solve_group(double theta_prev, double lo, double hi,
                                                    const HrfGrid &grid, ThetaScratch &sc, int q,
                                                    int *n_eval) {
    enum { INIT = 0, BRACKET = 1, NEWTON = 2, DONE = 3 };
    const int max_iter = 100;
    int phase = INIT, it = 0, evals = 0;
    double query = fmin(fmax(theta_prev, lo), hi);
    double result = query;
    double a = 0, ga = 0, ca = 0, step = 0, direction = 0, bound = 0;
    double xl = 0, xh = 0, x = 0, gx = 0, cx = 0, dx_old = 0, dx = 0;

    for (int guard = 0; guard < 2 * max_iter + 4; ++guard) {
        double g, c;
        theta_eval_group<G>(query, grid, sc, q, g, c);
        bool advance = false;   // NEWTON: pick the next point from (x, gx, cx, bracket)
        if (phase != DONE) ++evals;
        if (phase == INIT) {
            if (g == 0.0 || !isfinite(g)) {
                phase = DONE;
            } else {
                direction = g > 0.0 ? -1.0 : 1.0;
                bound = direction > 0.0 ? hi : lo;
                if (query == bound) {
                    phase = DONE;
                } else {
                    a = query;
                    ga = g;
                    ca = c;
                    step = ca > 0.0 ? fabs(ga / ca) : 0.125 * (hi - lo);
                    step = fmin(fmax(step, 1.0e-6), 0.25 * (hi - lo));
                    it = 0;
                    phase = BRACKET;
                    const double cand = a + direction * step;
                    query = direction > 0.0 ? fmin(cand, hi) : fmax(cand, lo);
                }
            }
        } else if (phase == BRACKET) {
            const double cand = query;
            if (g == 0.0) {
                result = cand;
                phase = DONE;
            } else if ((g > 0.0) != (ga > 0.0)) {
                if (ga < 0.0) {
                    xl = a;
                    xh = cand;
                } else {
                    xl = cand;
                    xh = a;
                }
                if (fabs(ga) < fabs(g)) {
                    x = a;
                    gx = ga;
                    cx = ca;
                } else {
                    x = cand;
                    gx = g;
                    cx = c;
                }
                dx_old = fabs(xh - xl);
                dx = dx_old;
                result = x;
                it = 0;
                phase = NEWTON;
                advance = true;
            } else {
                a = cand;
                ga = g;
                ca = c;
                if (cand == bound) {
                    result = bound;
                    phase = DONE;
                } else {
                    step *= 2.0;
                    if (++it >= max_iter) {
                        result = a;
                        phase = DONE;
                    } else {
                        const double nxt = a + direction * step;
                        query = direction > 0.0 ? fmin(nxt, hi) : fmax(nxt, lo);
                    }
                }
            }
        } else if (phase == NEWTON) {
            gx = g;
            cx = c;
            if (gx == 0.0) {
                phase = DONE;
            } else {
                if (gx < 0.0) xl = x; else xh = x;
                if (++it >= max_iter) phase = DONE; else advance = true;
            }
        }
        if (advance) {
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
            if (x_new == x) {
                phase = DONE;
            } else {
                x = x_new;
                result = x;
                if (fabs(dx) <= (newton_ok ? PB_THETA_XTOL_NEWTON : PB_THETA_XTOL_BISECT) * fmax(1.0, fabs(x)))
                    phase = DONE;
                else
                    query = x;
            }
        }
        if (phase == DONE) query = result;
        if (!__any_sync(PB_FULL, phase != DONE)) break;
    }
    if (n_eval) *n_eval = evals;
    return result;
}

}  // namespace pb
