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

// The outer phases are real function calls (not inlined into the solver kernels): ptxas then allocates the
// registers of the inner loop without the FP64 temporaries of these phases in the picture (inlined, every edit
// here re-shuffled the inner loop's allocation and its operand-reuse pattern, profiles/r01_rf_bandwidth.txt),
// and the state that lives across a call is saved once per outer iteration, i.e. once per nb_iter inner ones.
#ifdef PB_OUTER_NOINLINE
#define PB_OUTER_FN __device__ __noinline__
#else
#define PB_OUTER_FN __device__ __forceinline__
#endif
#ifndef PB_EVAL_SPLIT
#define PB_EVAL_SPLIT 0     /* 1: two accumulators per dot product in theta_eval_group (spills into the inner loop) */
#endif

namespace pb {

template <int G>
__device__ __forceinline__ double group_sum_f64(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(PB_FULL, v, o, G);
    return v;
}

template <int G>
__device__ __forceinline__ void hrf_eval_group_inl(double theta, const HrfGrid &grid, ThetaScratch &sc,
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

template <int G>
PB_OUTER_FN void hrf_eval_group(double theta, HrfGrid grid, ThetaScratch sc, int q) {
    hrf_eval_group_inl<G>(theta, grid, sc, q);
}

// same formula as frob_lipschitz_warp (pb_device.cuh), lanes of the group stride over the lags.
// These phases are latency bound (dependent FP64 chains, one warp instruction every ~14 cycles in the
// round-1 profile, 20 % of the warp time of the cfg3 kernel for 5.5 % of its instructions), so the loops
// below keep several independent chains in flight: a lane works on its two lags (d = q and d = q + G)
// in the same loop instead of one after the other, and plain sums are split over four accumulators.
template <int G>
PB_OUTER_FN double frob_lipschitz_group(ThetaScratch sc, int K, int T, int q) {
    // cs = cumsum(h), Ss = cumsum(cs): Ss[m] = sum_j (m - j + 1) h[j]; every lane sums its own entries
    for (int m = q; m < K; m += G) {
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = 0;
        for (; j + 3 <= m; j += 4) {
            const double h0 = sc.hs[j], h1 = sc.hs[j + 1], h2 = sc.hs[j + 2], h3 = sc.hs[j + 3];
            c0 += h0;
            c1 += h1;
            c2 += h2;
            c3 += h3;
            s0 = fma((double)(m - j + 1), h0, s0);
            s1 = fma((double)(m - j), h1, s1);
            s2 = fma((double)(m - j - 1), h2, s2);
            s3 = fma((double)(m - j - 2), h3, s3);
        }
        for (; j <= m; ++j) {
            const double h0 = sc.hs[j];
            c0 += h0;
            s0 = fma((double)(m - j + 1), h0, s0);
        }
        sc.cs[m] = (c0 + c1) + (c2 + c3);
        sc.Ss[m] = (s0 + s1) + (s2 + s3);
    }
    __syncwarp();
    const int E = K - 1;
    const double C = sc.cs[K - 1];
    double total = 0.0;
    for (int d0 = q; d0 < E && d0 < T; d0 += 2 * G) {
        // two lags per pass: d0 and d1 = d0 + G (independent chains)
        const int d1 = d0 + G;
        const bool on1 = d1 < E && d1 < T;
        double acc0 = 0.0, ss0 = 0.0, acc1 = 0.0, ss1 = 0.0;
        const int n0 = min(E - 1, T - 1 - d0), n1 = on1 ? min(E - 1, T - 1 - d1) : -1;
        const int nn = n0 > n1 ? n0 : n1;
        for (int m = 0; m <= nn; ++m) {
            const double cm = sc.cs[m];
            const int m0 = m + d0, m1 = m + d1;
            if (m <= n0) {
                acc0 = fma(cm, sc.cs[m0 < K - 1 ? m0 : K - 1], acc0);
                ss0 = fma(acc0, acc0, ss0);
            }
            if (m <= n1) {
                acc1 = fma(cm, sc.cs[m1 < K - 1 ? m1 : K - 1], acc1);
                ss1 = fma(acc1, acc1, ss1);
            }
        }
        const double e = C * C;
        const int Q0 = T - d0 - E, Q1 = T - d1 - E;
        if (Q0 > 0) {
            const double a = acc0, qq = (double)Q0;
            ss0 += qq * a * a + a * e * qq * (qq + 1.0) + e * e * qq * (qq + 1.0) * (2.0 * qq + 1.0) / 6.0;
        }
        if (on1 && Q1 > 0) {
            const double a = acc1, qq = (double)Q1;
            ss1 += qq * a * a + a * e * qq * (qq + 1.0) + e * e * qq * (qq + 1.0) * (2.0 * qq + 1.0) / 6.0;
        }
        total += (d0 == 0 ? 1.0 : 2.0) * ss0;
        if (on1) total += 2.0 * ss1;
    }
    const double SE = E > 0 ? sc.Ss[E - 1] : 0.0;
    const double dup = (E == 0) ? 1.0 : 0.0;
    double tt[4] = {0.0, 0.0, 0.0, 0.0};
    for (int n = q; n < T - E; n += 4 * G) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int nu = n + u * G;
            if (nu < T - E) {
                const double St = nu < E ? sc.Ss[nu] : SE + C * (double)(nu - E + 1);
                const double wgt = 2.0 * (double)(T - nu - E) - dup;
                const double v = C * St;
                tt[u] = fma(wgt * v, v, tt[u]);
            }
        }
    }
    const double t0 = tt[0] + tt[2], t1 = tt[1] + tt[3];
    total = group_sum_f64<G>(total + (t0 + t1));
    __syncwarp();
    return sqrt(total);
}

template <int G>
PB_OUTER_FN void gram_build_group(ThetaScratch sc, int K, int q) {
    // diagonals d0 = q and d1 = q + G in the same loop (two independent recurrences)
    for (int d0 = q; d0 < K; d0 += 2 * G) {
        const int d1 = d0 + G;
        const bool on1 = d1 < K;
        double acc0 = sc.Rz[d0], acc1 = on1 ? sc.Rz[d1] : 0.0;
        sc.M[d0] = acc0;
        sc.M[d0 * sc.KS] = acc0;
        if (on1) {
            sc.M[d1] = acc1;
            sc.M[d1 * sc.KS] = acc1;
        }
        for (int n = 1; n + d0 < K; ++n) {
            const double ze = sc.zend[n - 1];
            acc0 = fma(-ze, sc.zend[n - 1 + d0], acc0);
            sc.M[n * sc.KS + n + d0] = acc0;
            sc.M[(n + d0) * sc.KS + n] = acc0;
            if (n + d1 < K) {
                acc1 = fma(-ze, sc.zend[n - 1 + d1], acc1);
                sc.M[n * sc.KS + n + d1] = acc1;
                sc.M[(n + d1) * sc.KS + n] = acc1;
            }
        }
    }
    __syncwarp();
}

template <int G>
__device__ __forceinline__ void theta_eval_group(double theta, const HrfGrid &grid, ThetaScratch &sc,
                                                 int q, double &g, double &c) {
    hrf_eval_group_inl<G>(theta, grid, sc, q);
    const int K = grid.K;
    double pg = 0.0, pc = 0.0;
    // rows a0 = q and a1 = q + G of M in the same loop, each dot product over two accumulators: eight
    // independent DFMA chains instead of two (these quadratic forms were 24 % of the outer-phase time)
    for (int a0 = q; a0 < K; a0 += 2 * G) {
        const int a1 = a0 + G;
        const bool on1 = a1 < K;
        const double *r0 = sc.M + a0 * sc.KS;
        const double *r1 = sc.M + (on1 ? a1 : a0) * sc.KS;
        double qa0 = -sc.b[a0], qa0b = 0.0, q10 = 0.0, q10b = 0.0;
        double qa1 = on1 ? -sc.b[a1] : 0.0, qa1b = 0.0, q11 = 0.0, q11b = 0.0;
        int bb = 0;
        for (; PB_EVAL_SPLIT && bb + 1 < K; bb += 2) {
            const double ha = sc.hs[bb], hb = sc.hs[bb + 1], da = sc.h1s[bb], db = sc.h1s[bb + 1];
            const double m0a = r0[bb], m0b = r0[bb + 1], m1a = r1[bb], m1b = r1[bb + 1];
            qa0 = fma(m0a, ha, qa0);
            qa0b = fma(m0b, hb, qa0b);
            q10 = fma(m0a, da, q10);
            q10b = fma(m0b, db, q10b);
            qa1 = fma(m1a, ha, qa1);
            qa1b = fma(m1b, hb, qa1b);
            q11 = fma(m1a, da, q11);
            q11b = fma(m1b, db, q11b);
        }
        for (; bb < K; ++bb) {
            const double ha = sc.hs[bb], da = sc.h1s[bb];
            qa0 = fma(r0[bb], ha, qa0);
            q10 = fma(r0[bb], da, q10);
            qa1 = fma(r1[bb], ha, qa1);
            q11 = fma(r1[bb], da, q11);
        }
        {
            const double qa = qa0 + qa0b, q1 = q10 + q10b;
            pg = fma(sc.h1s[a0], qa, pg);
            pc = fma(sc.h1s[a0], q1, pc);
            pc = fma(sc.h2s[a0], qa, pc);
        }
        if (on1) {
            const double qa = qa1 + qa1b, q1 = q11 + q11b;
            pg = fma(sc.h1s[a1], qa, pg);
            pc = fma(sc.h1s[a1], q1, pc);
            pc = fma(sc.h2s[a1], qa, pc);
        }
    }
    g = group_sum_f64<G>(pg);
    c = group_sum_f64<G>(pc);
    __syncwarp();
}

template <int G>
PB_OUTER_FN double theta_solve_group(double theta_prev, double lo, double hi, HrfGrid grid, ThetaScratch sc,
                                     int q, int *n_eval) {
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
