// Register-tiled persistent bd kernel, CTA variant: ONE CTA OF NW WARPS PER VOXEL, for long series
// (T = 1200 at cfg4).  Thread t of the CTA owns the R contiguous samples [t R, t R + R) in
// registers (32 NW R >= T, R a multiple of 4).
//
// Why: with one warp per voxel a 1200-sample voxel needs R = 38-40 samples per lane, i.e. ~255
// registers, 8 resident warps per SM and a fully unrolled loop body of ~44 KB that no longer fits the
// instruction cache (measured IPC 0.52).  Two warps per voxel halve R: ~130 registers, 16 warps per
// SM and a 22 KB loop body.
//
// Differences from pb_fast.cuh / pb_fastg.cuh:
//  * the K-1 halo samples are exchanged through shared memory: every thread stores its R values as
//    float4 and loads the 28 preceding (or following) values as 7 float4.  Zero pads before and after
//    the voxel make the boundary handling free (no select), and this costs 12 instead of 54
//    instructions per convolution;
//  * the scans add a cross-warp carry (warp totals through shared memory);
//  * 4 __syncthreads per iteration (w visible, forward totals, residual visible, reverse totals);
//  * the Lipschitz constant and the theta solve are run by warp 0 with the code of pb_device.cuh.
// No early stopping (those calls go to pb_fast.cuh).  Tap 0 of the SPM HRF is identically zero.
//
// Reference code replaced: pybold/bold_signal.py:242-278, :281-382.
#pragma once
#include "pb_fast_registry.h"
#include "pb_generic.cuh"

#include "pb_tile.cuh"

// source order of the unrolled tile (pb_tile.cuh), picked with the register-file model on
// fast_bdc_kernel<float, 20, 28, 2, 6> (cfg4: 1587 -> 1423 modelled cycles per inner iteration; tools/order_search.py cta)
#ifndef PB_C_CONV_JDESC
#define PB_C_CONV_JDESC 1
#endif
#ifndef PB_C_CONV_RDESC
#define PB_C_CONV_RDESC 0
#endif
#ifndef PB_C_CONV_RB
#define PB_C_CONV_RB 5
#endif
#ifndef PB_C_CONV_DS
#define PB_C_CONV_DS 0
#endif
#ifndef PB_C_CORR_JDESC
#define PB_C_CORR_JDESC 1
#endif
#ifndef PB_C_CORR_RDESC
#define PB_C_CORR_RDESC 1
#endif
#ifndef PB_C_CORR_RB
#define PB_C_CORR_RB 5
#endif
#ifndef PB_C_CORR_DS
#define PB_C_CORR_DS 0
#endif

namespace pb {

template <typename real, int R, int KMAX, int NW>
struct CtaLayout {
    static constexpr int NT = NW * 32;
    static constexpr int SLOTS = NT * R;
    static constexpr int PAD = (KMAX + 3) & ~3;               // >= KMAX - 1, keeps float4 alignment
    // R = 8 / 16: the per-thread stride of R floats is a multiple of 8 banks and the 16-byte stores / loads
    // of a quarter warp would collide (round 1 measured R = 16 at 29 instead of 42 Tflop/s per slot and did
    // not build it).  SKEW inserts 4 unused floats after every 32 samples: sample i lives at
    // i + 4 (i >> 5), a thread's block and every aligned vector stay contiguous, and consecutive threads'
    // vectors fall into distinct banks again.  (i >> 5 is an arithmetic shift: the zero pad before the
    // series, i in [-PAD, -1], lands at i - 4.)
    static constexpr bool SKEW = R % 8 == 0;
    static constexpr int PADB = SKEW ? PAD + 8 : PAD;         // zero pad before sample 0
    __host__ __device__ static constexpr int pos(int i) { return SKEW ? i + ((i >> 5) << 2) : i; }
    static constexpr int BUF = PADB + pos(SLOTS + PAD) + 4;
    static constexpr int RED = NW * (2 * KMAX + 4);           // doubles for CTA-wide reductions
    __host__ __device__ static constexpr size_t bytes(int nb_iter) {
        return (((size_t)nb_iter * sizeof(real) + 15) & ~(size_t)15) +
               ((size_t)pb_scratch_doubles(KMAX) + RED + 2 * NW + 4) * sizeof(double) +
               (size_t)2 * BUF * sizeof(real);
    }
};

template <typename real, int R, int KMAX, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
fast_bdc_kernel(BdArgs<real> p) {
    static_assert(R % 4 == 0, "R must be a multiple of 4 (vector halo exchange)");
    using L = CtaLayout<real, R, KMAX, NW>;
    struct alignas(4 * sizeof(real)) V4 { real t[4]; };
    constexpr int NH = (KMAX - 1 + 3) / 4;                    // float4 per halo
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    real *beta = reinterpret_cast<real *>(smem);
    const size_t beta_bytes = ((size_t)p.nb_iter * sizeof(real) + 15) & ~(size_t)15;
    fill_momentum_table(beta, p.nb_iter);
    double *dbl = reinterpret_cast<double *>(smem + beta_bytes);
    ThetaScratch sc;
    sc.bind(dbl, KMAX);
    double *red = dbl + pb_scratch_doubles(KMAX);             // [NW][2 KMAX + 4]
    double *totf = red + L::RED;                              // [NW] forward warp totals
    double *totr = totf + NW;                                 // [NW] reverse warp totals
    double *bcast = totr + NW;                                // [4]  theta, Lipschitz, ...
    real *bufA = reinterpret_cast<real *>(bcast + 4);         // w / z with zero pads
    real *bufB = bufA + L::BUF;                               // residual with zero pads
    for (int i = tid; i < 2 * L::BUF; i += L::NT) bufA[i] = real(0);   // zero pads (and everything else once)
    real *const baseA = bufA + L::PADB, *const baseB = bufB + L::PADB;  // sample 0 of the two buffers
    real *myA = baseA + L::pos(tid * R), *myB = baseB + L::pos(tid * R);
    const int T = p.T, K = p.K, ntr = p.nb_iter + 2;
    const int i0 = tid * R;
    const int nvalid = max(0, min(R, T - i0));
    real *totf_r = reinterpret_cast<real *>(totf), *totr_r = reinterpret_cast<real *>(totr);
    __syncthreads();

    real w[R], dy[R], h[KMAX];

    // ---- small helpers over register tiles -------------------------------------------------
    auto put = [&](real *dst, const real (&a)[R]) {
#pragma unroll
        for (int r4 = 0; r4 < R / 4; ++r4) {
            V4 t;
#pragma unroll
            for (int e = 0; e < 4; ++e) t.t[e] = a[4 * r4 + e];
            reinterpret_cast<V4 *>(dst)[r4] = t;
        }
    };
    // hal[m-1] = value at voxel index (i0 - m), m = 1..4 NH
    auto halo_before = [&](const real *mine, real (&hal)[4 * NH]) {
#pragma unroll
        for (int c = 0; c < NH; ++c) {
            const V4 t = L::SKEW ? *reinterpret_cast<const V4 *>(mine - L::pos(i0) + L::pos(i0 - 4 * (c + 1)))
                                 : reinterpret_cast<const V4 *>(mine)[-1 - c];
#pragma unroll
            for (int e = 0; e < 4; ++e) hal[4 * c + (3 - e)] = t.t[e];
        }
    };
    // hal[k] = value at voxel index (i0 + R + k), k = 0..4 NH - 1
    auto halo_after = [&](const real *mine, real (&hal)[4 * NH]) {
#pragma unroll
        for (int c = 0; c < NH; ++c) {
            const V4 t = L::SKEW ? *reinterpret_cast<const V4 *>(mine - L::pos(i0) + L::pos(i0 + R + 4 * c))
                                 : reinterpret_cast<const V4 *>(mine + R)[c];
#pragma unroll
            for (int e = 0; e < 4; ++e) hal[4 * c + e] = t.t[e];
        }
    };
    auto conv_acc = [&](const real (&a)[R], const real (&hal)[4 * NH], real (&acc)[R]) {
        tile_conv<real, R, KMAX, 4 * NH, 1, PB_C_CONV_JDESC, PB_C_CONV_RDESC, PB_C_CONV_RB, PB_C_CONV_DS>(h, a, hal, acc);
    };
    auto corr_acc = [&](const real (&a)[R], const real (&hal)[4 * NH], real (&acc)[R]) {
        tile_corr<real, R, KMAX, 4 * NH, 1, PB_C_CORR_JDESC, PB_C_CORR_RDESC, PB_C_CORR_RB, PB_C_CORR_DS>(h, a, hal, acc);
    };
    auto mask_tail = [&](real (&a)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = r < nvalid ? a[r] : real(0);
    };
    // CTA-wide inclusive prefix of a register tile (one barrier inside)
    auto scan_fwd_cta = [&](real (&a)[R]) {
#pragma unroll
        for (int r = 1; r < R; ++r) a[r] += a[r - 1];
        real inc = a[R - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const real t = __shfl_up_sync(PB_FULL, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) totf_r[warp] = inc;
        real carry = __shfl_up_sync(PB_FULL, inc, 1);
        if (lane == 0) carry = real(0);
        __syncthreads();
#pragma unroll
        for (int ww = 0; ww < NW - 1; ++ww)
            if (ww < warp) carry += totf_r[ww];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] += carry;
    };
    auto scan_rev_cta = [&](real (&a)[R]) {
#pragma unroll
        for (int r = R - 2; r >= 0; --r) a[r] += a[r + 1];
        real inc = a[0];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const real t = __shfl_down_sync(PB_FULL, inc, d);
            if (lane + d < 32) inc += t;
        }
        if (lane == 0) totr_r[warp] = inc;
        real carry = __shfl_down_sync(PB_FULL, inc, 1);
        if (lane == 31) carry = real(0);
        __syncthreads();
#pragma unroll
        for (int ww = 1; ww < NW; ++ww)
            if (ww > warp) carry += totr_r[ww];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] += carry;
    };
    // CTA-wide sums of two doubles (two barriers; result on every thread)
    auto cta_sum2 = [&](double &a, double &b) {
        a = warp_sum(a);
        b = warp_sum(b);
        __syncthreads();
        if (lane == 0) {
            red[warp * (2 * KMAX + 4)] = a;
            red[warp * (2 * KMAX + 4) + 1] = b;
        }
        __syncthreads();
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) {
            sa += red[ww * (2 * KMAX + 4)];
            sb += red[ww * (2 * KMAX + 4) + 1];
        }
        a = sa;
        b = sb;
    };
    auto load_taps = [&]() {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) h[j] = j < K ? (real)sc.hs[j] : real(0);
    };

    // the first voxel of a CTA is static, the following ones come from the work queue when there is one
    __shared__ unsigned int next_task;
    for (int64_t v = blockIdx.x; v < p.V;) {
        const real *yv = p.y + v * T;
        real y[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r;
            y[r] = i < T ? yv[i] : real(0);
            const real prv = (i > 0 && i - 1 < T) ? yv[i - 1] : real(0);
            dy[r] = i < T ? y[r] - prv : real(0);
        }
        const double lam = (double)p.lbda[v * p.lbda_stride];
        double theta = (double)p.theta0[v * p.theta0_stride];
        __syncthreads();
        if (warp == 0) hrf_eval_warp(theta, p.grid, sc, lane);   // bold_signal.py:292 (theta_0: Q9)
        __syncthreads();
        load_taps();
        double r0, g0 = 0.0;
        if (p.z0) {                                               // bold_signal.py:298-301
            const real *zv = p.z0 + v * T;
            real z[R], hal[4 * NH], xr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = i0 + r;
                z[r] = i < T ? zv[i] : real(0);
                w[r] = (i > 0 && i < T) ? zv[i] - zv[i - 1] : real(0);
                xr[r] = -y[r];
            }
            put(myA, z);
            __syncthreads();
            halo_before(myA, hal);
            conv_acc(z, hal, xr);
            mask_tail(xr);
            real s2 = 0, s1 = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                s2 = fma(xr[r], xr[r], s2);
                s1 += fabs(w[r]);
            }
            r0 = (double)s2;
            g0 = (double)s1;
            cta_sum2(r0, g0);
        } else {
            real s2 = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                w[r] = real(0);
                s2 = fma(y[r], y[r], s2);
            }
            r0 = (double)s2;
            double zero = 0.0;
            cta_sum2(r0, zero);
        }
        const double j0 = r0 + lam * g0;
        real *Jv = p.out_J + v * (int64_t)ntr, *rv = p.out_r + v * (int64_t)ntr,
             *gv = p.out_g + v * (int64_t)ntr;
        if (tid == 0) {
            Jv[0] = real(1);
            rv[0] = real(1);
            gv[0] = (real)g0;
        }
        for (int idx = 0; idx <= p.nb_iter; ++idx) {
            const bool last = idx == p.nb_iter;                   // final deconvolution, :365-376
            __syncthreads();
            if (warp == 0) {
                const double Lw = frob_lipschitz_warp(sc, K, T, lane);
                if (lane == 0) bcast[1] = Lw;
            }
            __syncthreads();
            const double Lc = bcast[1];
            const real step = (real)(1.0 / Lc), th = (real)(lam / Lc);
            for (int j = 0; j < p.nb_iter; ++j) {                 // _loops_deconv, :259-276
                real hal[4 * NH], res[R], g[R];
                put(myA, w);
                __syncthreads();                                  // B1: w visible
                halo_before(myA, hal);
#pragma unroll
                for (int r = 0; r < R; ++r) res[r] = -dy[r];
                conv_acc(w, hal, res);
                scan_fwd_cta(res);                                // B2 inside
                mask_tail(res);
                put(myB, res);
                __syncthreads();                                  // B3: residual visible
                halo_after(myB, hal);
#pragma unroll
                for (int r = 0; r < R; ++r) g[r] = real(0);
                corr_acc(res, hal, g);
                scan_rev_cta(g);                                  // B4 inside
                const real ob = real(1) + beta[j];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const real u = fma(-step, g[r], w[r]);
                    const real cl = fmin(fmax(u, -th), th);
                    w[r] = fma(-ob, cl, u);
                }
            }
            real z[R], hal[4 * NH];
#pragma unroll
            for (int r = 0; r < R; ++r) z[r] = w[r];
            scan_fwd_cta(z);
            put(myA, z);
            __syncthreads();
            halo_before(myA, hal);
            if (!last) {
                // ---- theta step (:329-334): b = Z^T y, Rz = autocorrelation of z ----
                real zm[R];
#pragma unroll
                for (int r = 0; r < R; ++r) zm[r] = z[r];
                mask_tail(zm);
#pragma unroll
                for (int a = 0; a < KMAX; ++a) {
                    real pb_ = 0, pr = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int id = r - a;
                        const real zz = id >= 0 ? z[id >= 0 ? id : 0] : hal[id >= 0 ? 0 : -id - 1];
                        pb_ = fma(y[r], zz, pb_);
                        pr = fma(zm[r], zz, pr);
                    }
                    pb_ = warp_sum(pb_);
                    pr = warp_sum(pr);
                    if (lane == 0) {
                        red[warp * (2 * KMAX + 4) + 4 + a] = (double)pb_;
                        red[warp * (2 * KMAX + 4) + 4 + KMAX + a] = (double)pr;
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int pidx = T - 1 - (i0 + r);
                    if (pidx >= 0 && pidx < K) sc.zend[pidx] = (double)z[r];
                }
                for (int a = tid; a < K; a += L::NT)
                    if (a >= T) sc.zend[a] = 0.0;
                __syncthreads();
                if (warp == 0) {
                    for (int a = lane; a < K; a += 32) {
                        double sb = 0.0, sr = 0.0;
#pragma unroll
                        for (int ww = 0; ww < NW; ++ww) {
                            sb += red[ww * (2 * KMAX + 4) + 4 + a];
                            sr += red[ww * (2 * KMAX + 4) + 4 + KMAX + a];
                        }
                        sc.b[a] = sb;
                        sc.Rz[a] = sr;
                    }
                    __syncwarp();
                    gram_build_warp(sc, K, lane);
                    const double th_n = theta_solve_warp(theta, p.theta_lo, p.theta_hi, p.grid, sc,
                                                         lane, nullptr);
                    hrf_eval_warp(th_n, p.grid, sc, lane);
                    if (lane == 0) bcast[0] = th_n;
                }
                __syncthreads();
                theta = bcast[0];
                load_taps();
            }
            // ---- cost trace: x = h * z with the (new) taps ----
            real xr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) xr[r] = -y[r];
            conv_acc(z, hal, xr);
            mask_tail(xr);
            real s2 = 0, s1 = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                s2 = fma(xr[r], xr[r], s2);
                s1 += fabs(w[r]);
            }
            double rr = (double)s2, gg = (double)s1;
            cta_sum2(rr, gg);
            const double eps = last ? 0.0 : 1.0e-30;
            if (tid == 0) {
                Jv[idx + 1] = (real)((rr + lam * gg) / j0 + eps);
                rv[idx + 1] = (real)(rr / r0 + eps);
                gv[idx + 1] = (real)gg;
            }
            if (last) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = i0 + r;
                    if (i < T) {
                        p.out_x[v * T + i] = xr[r] + y[r];
                        p.out_z[v * T + i] = z[r];
                        p.out_dz[v * T + i] = w[r];
                    }
                }
            }
        }
        __syncthreads();
        for (int a = tid; a < K; a += L::NT) p.out_h[v * K + a] = (real)sc.hs[a];
        if (tid == 0) {
            p.out_theta[v] = (real)theta;
            p.out_ntrace[v] = ntr;
            if (p.queue) next_task = atomicAdd(p.queue, 1u);
        }
        __syncthreads();
        v = p.queue ? (int64_t)gridDim.x + (int64_t)next_task : v + gridDim.x;
    }
}

template <int R, int KMAX, int NW>
bool fastc_shape_ok(int T, int K) {
    // the last used thread may be partial, threads beyond it idle; up to half of the threads may idle
    // (the dispatcher picks the variant with the fewest slots among those that match)
    // (the smallest variants of the many-tap families, one warp of R = 4, also take everything shorter;
    // with K <= 32 short series are better served by the group / warp kernels)
    // the many-tap families (KMAX >= 40) only serve K > 28: with fewer taps the K <= 28 variants, the group
    // and the warp kernels are the better choice
    return K <= KMAX && (KMAX <= 32 || K > 28) && T <= NW * 32 * R &&
           (T > NW * 16 * R - R || (NW == 1 && R == 4)) && T >= 1;
}

template <typename real, int R, int KMAX, int NW, int MINB>
int fast_bdc_launch(const BdArgs<real> &a, cudaStream_t stream) {
    using L = CtaLayout<real, R, KMAX, NW>;
    const size_t smem = L::bytes(a.nb_iter);
    auto kern = fast_bdc_kernel<real, R, KMAX, NW, MINB>;
    int dev = 0, sms = 0, max_smem = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_smem) return FAST_NO_MATCH;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return FAST_NO_MATCH;
    const int64_t cap = (int64_t)sms * occ;
    const int grid = (int)(a.V < cap ? a.V : cap);
    if (a.queue) {
        e = cudaMemsetAsync(a.queue, 0, sizeof(unsigned int), stream);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, NW * 32, smem, stream>>>(a);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <typename real, int R, int KMAX, int NW, int MINB>
int fast_bdc_wave(int nb_iter) {
    using L = CtaLayout<real, R, KMAX, NW>;
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = fast_bdc_kernel<real, R, KMAX, NW, MINB>;
    const size_t smem = L::bytes(nb_iter);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return sms * occ;
}

}  // namespace pb
