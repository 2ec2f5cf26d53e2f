// Row N3 of SURVEY.md 8(f): on-device synthetic voxels and the post-processing of the ICASSP-2019
// simulation (examples/icassp_2019/simulation.py:27-48 generation pattern, :143-147 error metrics,
// pybold/utils.py:112-138 inf_norm), so that a simulation sweep never leaves the GPU.
//
// The generator is this package's own ("philox-v1", restated in NumPy in pybold_b200/synth.py):
// the reference's pybold/data.py draws from NumPy's global Mersenne twister with a rejection loop
// and is broken on NumPy >= 1.24.  Every number depends only on
// (seed, global voxel index, sample index) through the counter-based Philox4x32-10 generator, so any
// voxel range can be produced on any rank:
//   key      = (seed & 0xffffffff, seed >> 32)
//   stream 0 = Philox(ctr = (q, 0, g_lo, g_hi)), q = 0, 1, ...: u[4q .. 4q+3]
//              delta   = lo + (hi - lo) * (u[0] + 0.5) / 2^32
//              onset_e = floor(u[1 + e] * n_onset / 2^32),  e < nb_events
//   stream 1 = Philox(ctr = (p, 1, g_lo, g_hi)) -> four Gaussians n[4p .. 4p+3] (Box-Muller on the pairs
//              (u0, u1), (u2, u3): sqrt(-2 ln a) * {cos, sin}(2 pi b), a, b = (u + 0.5) / 2^32)
//   z        = sum of nb_events unit boxcars of `blk` samples at the onsets
//   x        = (h / max|h|) * z (causal, truncated), h = SPM taps at dilation delta (hrf_model.py:12-39)
//   y        = x + n * ||x|| / (||n|| + eps) / 10^(snr_db / 20)        (rule of pybold/data.py:439-444)
// All arithmetic in double; only the stores round to `real`.
#pragma once
#include "pb_device.cuh"

namespace pb {

struct Philox4 { uint32_t v[4]; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r > 0) {
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

__device__ __forceinline__ double u01(uint32_t u) { return ((double)u + 0.5) * (1.0 / 4294967296.0); }

__device__ __forceinline__ void gauss4(uint32_t p, uint32_t g_lo, uint32_t g_hi, uint32_t k0, uint32_t k1,
                                       double (&n)[4]) {
    const Philox4 r = philox4x32_10(p, 1u, g_lo, g_hi, k0, k1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const double rad = sqrt(-2.0 * log(u01(r.v[2 * h])));
        double sn, cs;
        sincospi(2.0 * u01(r.v[2 * h + 1]), &sn, &cs);
        n[2 * h] = rad * cs;
        n[2 * h + 1] = rad * sn;
    }
}

struct SynthArgs {
    uint64_t seed;
    int64_t first_voxel;
    HrfGrid grid;
    double delta_lo, delta_hi, snr_db;
    int nb_events, blk;
    int64_t V;
    int T;
};

constexpr int PB_SYNTH_MAX_EVENTS = 31;

// one warp per voxel; shared memory per warp: z[T] floats (small integers, exact) + h[K] doubles
template <typename real>
__global__ void synth_voxels_kernel(SynthArgs a, real *out_y, real *out_z, real *out_delta) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int T = a.T, K = a.grid.K;
    const size_t warp_bytes = (((size_t)T * sizeof(float) + 15) & ~(size_t)15) + (size_t)K * sizeof(double);
    float *zs = reinterpret_cast<float *>(smem + (size_t)warp * warp_bytes);
    double *hs = reinterpret_cast<double *>(smem + (size_t)warp * warp_bytes +
                                            (((size_t)T * sizeof(float) + 15) & ~(size_t)15));
    const uint32_t k0 = (uint32_t)(a.seed & 0xffffffffu), k1 = (uint32_t)(a.seed >> 32);
    const int n_onset = max(T - a.blk - 1, 1);
    for (int64_t v = (int64_t)blockIdx.x * nw + warp; v < a.V; v += (int64_t)gridDim.x * nw) {
        const uint64_t g = (uint64_t)(a.first_voxel + v);
        const uint32_t g_lo = (uint32_t)(g & 0xffffffffu), g_hi = (uint32_t)(g >> 32);
        // ---- paradigm: lane e < nb_events owns onset e; u[0] is the dilation ----
        const int slot = lane + 1;                                  // index into the u stream
        const Philox4 mine = philox4x32_10((uint32_t)(slot >> 2), 0u, g_lo, g_hi, k0, k1);
        const uint32_t u_mine = mine.v[slot & 3];
        const int onset = (int)(((uint64_t)u_mine * (uint64_t)n_onset) >> 32);
        const Philox4 first = philox4x32_10(0u, 0u, g_lo, g_hi, k0, k1);
        const double delta = a.delta_lo + (a.delta_hi - a.delta_lo) * u01(first.v[0]);
        for (int i0 = 0; i0 < T; i0 += 32) {      // warp-uniform trip count: the shuffles need all lanes
            const int i = i0 + lane;
            int cnt = 0;
            for (int e = 0; e < a.nb_events; ++e) {
                const int o = __shfl_sync(PB_FULL, onset, e);
                cnt += (i >= o && i < o + a.blk) ? 1 : 0;
            }
            if (i < T) zs[i] = (float)cnt;
        }
        // ---- normalised taps ----
        double hmax = 0.0;
        for (int j = lane; j < K; j += 32) {
            const double hv = hrf_value(delta, a.grid.t(j));
            hs[j] = hv;
            hmax = fmax(hmax, fabs(hv));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hmax = fmax(hmax, __shfl_xor_sync(PB_FULL, hmax, o));
        __syncwarp();
        const double hinv = 1.0 / hmax;
        // ---- pass 1: norms of the clean signal and of the noise ----
        double sx = 0.0, sn = 0.0;
        const int ngrp = (T + 3) >> 2;
        for (int p = lane; p < ngrp; p += 32) {
            double n[4];
            gauss4((uint32_t)p, g_lo, g_hi, k0, k1, n);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * p + e;
                if (i < T) {
                    double x = 0.0;
                    const int jm = min(K - 1, i);
                    for (int j = 0; j <= jm; ++j) x = fma(hs[j], (double)zs[i - j], x);
                    x *= hinv;
                    sx = fma(x, x, sx);
                    sn = fma(n[e], n[e], sn);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(PB_FULL, sx, o);
            sn += __shfl_xor_sync(PB_FULL, sn, o);
        }
        const double scale = sqrt(sx) / (sqrt(sn) + 2.220446049250313e-16) / exp10(a.snr_db / 20.0);
        // ---- pass 2: y = x + scale n (the noise is regenerated, nothing is staged) ----
        for (int p = lane; p < ngrp; p += 32) {
            double n[4];
            gauss4((uint32_t)p, g_lo, g_hi, k0, k1, n);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * p + e;
                if (i < T) {
                    double x = 0.0;
                    const int jm = min(K - 1, i);
                    for (int j = 0; j <= jm; ++j) x = fma(hs[j], (double)zs[i - j], x);
                    out_y[v * T + i] = (real)fma(scale, n[e], x * hinv);
                    if (out_z) out_z[v * T + i] = (real)zs[i];
                }
            }
        }
        if (out_delta && lane == 0) out_delta[v] = (real)delta;
        __syncwarp();
    }
}

// out[v, :] = x[v, :] / (max |x[v, :]| + 1e-12)        (pybold/utils.py:112-138, 2-D input, axis = 1)
template <typename real>
__global__ void inf_norm_kernel(const real *__restrict__ x, real *out, int64_t V, int T) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < V; v += nwarp) {
        const real *row = x + v * T;
        real m = 0;
        for (int i = lane; i < T; i += 32) m = fmax(m, fabs(row[i]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(PB_FULL, m, o));
        const real d = m + (real)1.0e-12;
        for (int i = lane; i < T; i += 32) out[v * T + i] = row[i] / d;
    }
}

// err[v] = ||est[v, :] - ref[v, :]|| / ||ref[v, :]||       (examples/icassp_2019/simulation.py:143-147);
// ref_stride = 0: one reference row for every voxel (the true HRF)
template <typename real>
__global__ void rel_l2_err_kernel(const real *__restrict__ est, const real *__restrict__ ref,
                                  int64_t ref_stride, real *out_err, int64_t V, int T) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < V; v += nwarp) {
        double sd = 0.0, sr = 0.0;
        for (int i = lane; i < T; i += 32) {
            const double r = (double)ref[v * ref_stride + i], d = (double)est[v * T + i] - r;
            sd = fma(d, d, sd);
            sr = fma(r, r, sr);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sd += __shfl_xor_sync(PB_FULL, sd, o);
            sr += __shfl_xor_sync(PB_FULL, sr, o);
        }
        if (lane == 0) out_err[v] = (real)(sqrt(sd) / sqrt(sr));
    }
}

// One outer iteration of the noise-constrained lambda loop of deconv(lbda=None)
// (pybold/bold_signal.py:139-162), fused over the batch: voxels still active take the new inner-loop
// result, then r = ||x - y||^2, g = ||diff_z||_1, alpha += mu (r - T sigma^2), lbda = 1 / (2 alpha).
// One warp per voxel, rows read once and (for active voxels) written once.
template <typename real>
__global__ void noise_step_kernel(const real *__restrict__ xn, const real *__restrict__ zn,
                                  const real *__restrict__ wn, const real *__restrict__ y,
                                  const real *__restrict__ sigma, const unsigned char *__restrict__ active,
                                  double mu, real *x, real *z, real *w, real *alpha, real *lbda,
                                  real *out_r, real *out_g, int64_t V, int T) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < V; v += nwarp) {
        const bool on = active[v] != 0;
        real sr = 0, sg = 0;
        for (int i = lane; i < T; i += 32) {
            const int64_t o = v * T + i;
            real xv, wv;
            if (on) {
                xv = xn[o];
                wv = wn[o];
                x[o] = xv;
                w[o] = wv;
                z[o] = zn[o];
            } else {
                xv = x[o];
                wv = w[o];
            }
            const real d = xv - y[o];
            sr = fma(d, d, sr);
            sg += fabs(wv);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            sr += __shfl_xor_sync(PB_FULL, sr, s);
            sg += __shfl_xor_sync(PB_FULL, sg, s);
        }
        if (lane == 0) {
            real a = alpha[v];
            if (on) {
                const real sgm = sigma[v];
                a = a + (real)mu * (sr - (real)T * sgm * sgm);
                alpha[v] = a;
            }
            lbda[v] = real(1) / (real(2) * a);
            out_r[v] = sr;
            out_g[v] = sg;
        }
    }
}

}  // namespace pb
