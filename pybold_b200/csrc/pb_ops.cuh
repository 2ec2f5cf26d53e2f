// Stand-alone operator kernels (A1, A2, A4, A5, A9, A10 of SURVEY.md section 8): the pieces of
// pybold/linear.py, pybold/convolution.py, pybold/hrf_model.py and pybold/utils.py that the
// solvers fuse, exposed one by one so that the reference's `op` / `adj` protocol and its
// operator tests have a device counterpart.  One warp per voxel row, row staged in shared
// memory (coalesced load / store, HBM bound).
#pragma once
#include "pb_generic.cuh"

namespace pb {

enum OpKind { OP_INTEG = 0, OP_INTEG_ADJ, OP_CONV, OP_CONV_ADJ, OP_HRFINTEG, OP_HRFINTEG_ADJ };

template <typename real, int OP>
__global__ void op_kernel(const real *h, int64_t h_stride, const real *x, real *out, int64_t V,
                          int T, int K, GenLayout lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double *scratch;
    GenVoxel<real> g = gen_bind<real>(smem + (size_t)warp * lay.warp_bytes(sizeof(real)), lay, T, K,
                                      lane, scratch);
    for (int64_t v = (int64_t)blockIdx.x * nwarp + warp; v < V; v += (int64_t)gridDim.x * nwarp) {
        for (int i = lane; i < T; i += 32) g.as[i] = x[v * T + i];
        if (OP >= OP_CONV)
            for (int a = lane; a < K; a += 32) g.hr[a] = h[v * h_stride + a];
        __syncwarp();
        real *res = g.as;
        if (OP == OP_INTEG) {
            g.scan_fwd(g.as);
        } else if (OP == OP_INTEG_ADJ) {
            g.scan_rev(g.as);
        } else if (OP == OP_CONV) {
            g.conv(g.as, nullptr, g.bs);
            res = g.bs;
        } else if (OP == OP_CONV_ADJ) {
            g.corr(g.as, g.bs);
            res = g.bs;
        } else if (OP == OP_HRFINTEG) {
            g.scan_fwd(g.as);
            g.conv(g.as, nullptr, g.bs);
            res = g.bs;
        } else {
            g.corr(g.as, g.bs);
            g.scan_rev(g.bs);
            res = g.bs;
        }
        for (int i = lane; i < T; i += 32) out[v * T + i] = res[i];
        __syncwarp();
    }
}

// spm_hrf (pybold/hrf_model.py:12-39): taps at the kept samples; the optional normalisation
// divides by max(hrf + 1e-30) over the reference's full 1 ms grid (N = int(dur / dt) points).
template <typename real>
__global__ void spm_hrf_kernel(const real *theta, HrfGrid grid, int n_fine, int normalized,
                               real *out_h, int64_t V) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int64_t v = (int64_t)blockIdx.x * nwarp + warp; v < V; v += (int64_t)gridDim.x * nwarp) {
        const double th = (double)theta[v];
        double scale = 1.0;
        if (normalized) {
            double mx = -1.0e300;
            for (int n = lane; n < n_fine; n += 32)
                mx = fmax(mx, hrf_value(th, (double)n * grid.t_step) + 1.0e-30);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(PB_FULL, mx, o));
            scale = mx;
        }
        for (int a = lane; a < grid.K; a += 32) {
            double hh, h1, h2;
            hrf_tap(th, grid.t(a), hh, h1, h2);
            out_h[v * grid.K + a] = (real)(normalized ? hh / scale : hh);
        }
    }
}

// spectral_radius_est (pybold/utils.py:94-109) with the start vector supplied.
template <typename real>
__global__ void lipschitz_power_kernel(const real *h, int64_t h_stride, const real *x0,
                                       int64_t x0_stride, int nb_iter, double tol, real *out_L,
                                       int64_t V, int T, int K, GenLayout lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double *scratch;
    GenVoxel<real> g = gen_bind<real>(smem + (size_t)warp * lay.warp_bytes(sizeof(real)), lay, T, K,
                                      lane, scratch);
    for (int64_t v = (int64_t)blockIdx.x * nwarp + warp; v < V; v += (int64_t)gridDim.x * nwarp) {
        for (int i = lane; i < T; i += 32) g.ws[i] = x0[v * x0_stride + i];
        for (int a = lane; a < K; a += 32) g.hr[a] = h[v * h_stride + a];
        __syncwarp();
        double n_old = sqrt(g.sumsq(g.ws));
        double n_new = n_old;
        for (int it = 0; it < nb_iter; ++it) {
            for (int i = lane; i < T; i += 32) g.as[i] = g.ws[i];
            __syncwarp();
            g.scan_fwd(g.as);
            g.conv(g.as, nullptr, g.bs);
            g.adjoint();
            const real inv = (real)(1.0 / n_old);
            for (int i = lane; i < T; i += 32) g.as[i] *= inv;
            __syncwarp();
            n_new = sqrt(g.sumsq(g.as));
            if (fabs(n_new - n_old) < tol) break;
            for (int i = lane; i < T; i += 32) g.ws[i] = g.as[i];
            __syncwarp();
            n_old = n_new;
        }
        if (lane == 0) out_L[v] = (real)n_new;
        __syncwarp();
    }
}

template <typename real>
__global__ void lipschitz_frob_kernel(const real *h, int64_t h_stride, real *out_L, int64_t V,
                                      int T, int K, int kp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    ThetaScratch sc;
    double *base = reinterpret_cast<double *>(smem) + (size_t)warp * 3 * kp;
    sc.hs = base;
    sc.cs = base + kp;
    sc.Ss = base + 2 * kp;
    for (int64_t v = (int64_t)blockIdx.x * nwarp + warp; v < V; v += (int64_t)gridDim.x * nwarp) {
        for (int a = lane; a < K; a += 32) sc.hs[a] = (double)h[v * h_stride + a];
        __syncwarp();
        const double Lc = frob_lipschitz_warp(sc, K, T, lane);
        if (lane == 0) out_L[v] = (real)Lc;
        __syncwarp();
    }
}

// Layout adapter (row N4 of SURVEY.md 8(f)): the ingest side of the reference pipeline holds voxel
// matrices time-major, [T, V] (NiftiMasker.fit_transform, examples/icassp_2019/validation.py:90-103,
// consumed as `voxels.T`); the solvers want [V, T].  32 x 32 tiles through padded shared memory:
// coalesced 128-byte rows on both sides, HBM bound (2 x 4 bytes per sample).
template <typename real>
__global__ void transpose_kernel(const real *in, real *out, int64_t rows, int64_t cols) {
    __shared__ real tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;      // 32 x 8 threads
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t r = r0 + ty + k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + k][tx] = in[r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t c = c0 + ty + k, r = r0 + tx;
        if (r < rows && c < cols) out[c * rows + r] = tile[tx][ty + k];
    }
}

// hrf_estim / the theta step alone (pybold/bold_signal.py:217-239, :329-334)
template <typename real>
__global__ void hrf_estim_kernel(const real *z, const real *y, HrfGrid grid, const real *theta0,
                                 int64_t theta0_stride, double lo, double hi, real *out_theta,
                                 real *out_h, real *out_cost, int64_t V, int T, GenLayout lay) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int K = grid.K;
    double *scratch;
    GenVoxel<real> g = gen_bind<real>(smem + (size_t)warp * lay.warp_bytes(sizeof(real)), lay, T, K,
                                      lane, scratch);
    ThetaScratch sc;
    sc.bind(scratch, lay.kp);
    for (int64_t v = (int64_t)blockIdx.x * nwarp + warp; v < V; v += (int64_t)gridDim.x * nwarp) {
        for (int i = lane; i < T; i += 32) {
            g.as[i] = z[v * T + i];
            g.ys[i] = y[v * T + i];
        }
        __syncwarp();
        g.theta_moments(sc);
        gram_build_warp(sc, K, lane);
        const double th = theta_solve_warp((double)theta0[v * theta0_stride], lo, hi, grid, sc,
                                           lane, nullptr);
        hrf_eval_warp(th, grid, sc, lane);
        g.load_taps(sc);
        g.conv(g.as, g.ys, g.bs);
        const double cost = 0.5 * g.sumsq(g.bs);
        for (int a = lane; a < K; a += 32) out_h[v * K + a] = (real)sc.hs[a];
        if (lane == 0) {
            out_theta[v] = (real)th;
            out_cost[v] = (real)cost;
        }
        __syncwarp();
    }
}

}  // namespace pb
