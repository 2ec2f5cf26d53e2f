// Register-resident row operators (A1, A2, A5 of SURVEY.md section 8): pybold/linear.py:15-43
// (DiscretInteg.op / adj), pybold/linear.py:73-113 (ConvAndLinear.op / adj) and
// pybold/convolution.py:135-196 (simple_convolve / simple_retro_convolve) on [V, T] batches.
//
// HBM-bound kernels: 8 bytes of traffic per sample (one read, one write).  One warp per row; the
// whole row is fetched with 16-byte loads into registers before anything else happens (all loads
// of a row in flight at once), scans run in registers (in-vector prefix + shuffle scan of the
// vector totals + running carry), the K-tap convolution reads 16-byte windows from a zero-framed
// shared-memory copy of the row with the taps in registers, and the result leaves with 16-byte
// stores.  Rows must start on 16-byte boundaries (T a multiple of 16 / sizeof(real)), T <= 32 *
// VEC * NCH and K <= KMAX; other shapes are served by op_kernel (pb_ops.cuh).
#pragma once
#include "pb_ops.cuh"

namespace pb {

template <typename real>
struct alignas(16) Vec16 {
    static constexpr int N = 16 / sizeof(real);
    real t[N];
};

template <typename real, int NCH>
struct RowRegs {
    static constexpr int VEC = Vec16<real>::N;
    Vec16<real> d[NCH];

    __device__ __forceinline__ void load(const real *__restrict__ row, int nvec, int lane) {
        const Vec16<real> *src = reinterpret_cast<const Vec16<real> *>(row);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i = c * 32 + lane;
            if (i < nvec) {
                d[c] = src[i];
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) d[c].t[e] = real(0);
            }
        }
    }
    __device__ __forceinline__ void store(real *row, int nvec, int lane) const {
        Vec16<real> *dst = reinterpret_cast<Vec16<real> *>(row);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i = c * 32 + lane;
            if (i < nvec) dst[i] = d[c];
        }
    }
    // cumsum along the row (pybold/linear.py:28)
    __device__ __forceinline__ void scan_fwd(int lane) {
        real carry = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
            for (int e = 1; e < VEC; ++e) d[c].t[e] += d[c].t[e - 1];
            real inc = d[c].t[VEC - 1];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const real t = __shfl_up_sync(PB_FULL, inc, s);
                if (lane >= s) inc += t;
            }
            real ex = __shfl_up_sync(PB_FULL, inc, 1);
            ex = (lane == 0 ? real(0) : ex) + carry;
#pragma unroll
            for (int e = 0; e < VEC; ++e) d[c].t[e] += ex;
            carry += __shfl_sync(PB_FULL, inc, 31);
        }
    }
    // reversed cumsum (pybold/linear.py:43); the zero fill beyond T contributes nothing
    __device__ __forceinline__ void scan_rev(int lane) {
        real carry = 0;
#pragma unroll
        for (int c = NCH - 1; c >= 0; --c) {
#pragma unroll
            for (int e = VEC - 2; e >= 0; --e) d[c].t[e] += d[c].t[e + 1];
            real inc = d[c].t[0];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const real t = __shfl_down_sync(PB_FULL, inc, s);
                if (lane + s < 32) inc += t;
            }
            real ex = __shfl_down_sync(PB_FULL, inc, 1);
            ex = (lane == 31 ? real(0) : ex) + carry;
#pragma unroll
            for (int e = 0; e < VEC; ++e) d[c].t[e] += ex;
            carry += __shfl_sync(PB_FULL, inc, 0);
        }
    }
};

template <typename real, bool REV, int NCH>
__global__ void __launch_bounds__(256)
rows_scan_kernel(const real *__restrict__ x, real *out, int64_t V, int T) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nvec = T / Vec16<real>::N;
    for (int64_t v = warp; v < V; v += nwarp) {
        RowRegs<real, NCH> r;
        r.load(x + v * T, nvec, lane);
        if (REV) r.scan_rev(lane); else r.scan_fwd(lane);
        r.store(out + v * T, nvec, lane);
    }
}

// K-tap causal convolution / anti-causal correlation of the rows, optionally fused with the
// integration operator: OP_CONV, OP_CONV_ADJ, OP_HRFINTEG (conv o cumsum), OP_HRFINTEG_ADJ
// (reversed cumsum o corr).
//
// Bound by shared-memory bandwidth, not HBM (ncu, K = 28: 70 % of the shared-memory wavefront peak,
// l1tex 87 %, at 0.71 of the HBM copy rate): every sample is read by (K + 6) / 4 lanes.  Measured
// and rejected: two adjacent output vectors per lane (half the window loads, but 32-byte lane
// strides double the wavefronts per load: 0.65); three CTAs per SM (spills, no gain).
template <int KMAX, int VEC>
__host__ __device__ constexpr int rows_conv_nc() { return (KMAX - 1 + VEC - 1) / VEC + 1; }

template <typename real, int KMAX, int NCH>
struct RowsConvLayout {
    static constexpr int VEC = Vec16<real>::N;
    static constexpr int PAD = VEC * (rows_conv_nc<KMAX, VEC>() - 1);   // zero frame on each side
    static constexpr int TP = 32 * VEC * NCH;
    static constexpr int ROW = 2 * PAD + TP;
    static constexpr size_t WARP_BYTES = (size_t)(ROW + KMAX) * sizeof(real);
};

template <typename real, int OP, int KMAX, int NCH>
__global__ void __launch_bounds__(256, 2)
rows_conv_kernel(const real *__restrict__ h, int64_t h_stride, const real *__restrict__ x, real *out,
                 int64_t V, int T, int K) {
    using L = RowsConvLayout<real, KMAX, NCH>;
    constexpr int VEC = L::VEC, NC = rows_conv_nc<KMAX, VEC>();
    constexpr bool ADJ = OP == OP_CONV_ADJ || OP == OP_HRFINTEG_ADJ;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    real *buf = reinterpret_cast<real *>(smem + (size_t)warp * L::WARP_BYTES);
    real *A = buf + L::PAD;              // first sample of the row
    real *hs = buf + L::ROW;             // this row's taps, zero beyond K
    for (int i = lane; i < L::PAD; i += 32) {
        buf[i] = real(0);
        buf[L::PAD + L::TP + i] = real(0);
    }
    const int nvec = T / VEC;
    real taps[KMAX];
    bool have_taps = false;
    for (int64_t v = (int64_t)blockIdx.x * nw + warp; v < V; v += (int64_t)gridDim.x * nw) {
        RowRegs<real, NCH> r;
        r.load(x + v * T, nvec, lane);
        if (h_stride != 0 || !have_taps) {
            __syncwarp();
            for (int a = lane; a < KMAX; a += 32) hs[a] = a < K ? h[v * h_stride + a] : real(0);
        }
        if (OP == OP_HRFINTEG) r.scan_fwd(lane);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            reinterpret_cast<Vec16<real> *>(A)[c * 32 + lane] = r.d[c];
        __syncwarp();
        if (h_stride != 0 || !have_taps) {
#pragma unroll
            for (int a = 0; a < KMAX; ++a) taps[a] = hs[a];
            have_taps = true;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i0 = (c * 32 + lane) * VEC;
            real acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = real(0);
            if (c * 32 * VEC < T) {          // warp-uniform: chunks past the row end stay zero
#pragma unroll
                for (int w = 0; w < NC; ++w) {
                    const Vec16<real> xv =
                        *reinterpret_cast<const Vec16<real> *>(A + (ADJ ? i0 + VEC * w : i0 - VEC * w));
#pragma unroll
                    for (int t = 0; t < VEC; ++t)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const int j = ADJ ? VEC * w + t - e : VEC * w + e - t;
                            if (j >= 0 && j < KMAX) acc[e] = fma(taps[j >= 0 && j < KMAX ? j : 0], xv.t[t], acc[e]);
                        }
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) r.d[c].t[e] = acc[e];
        }
        if (OP == OP_HRFINTEG_ADJ) {
            // samples beyond T of the last vector chunk are zero inputs, but the correlation of the
            // zero fill is zero as well: nothing to mask before the reversed scan
            r.scan_rev(lane);
        }
        r.store(out + v * T, nvec, lane);
        __syncwarp();                        // the next row overwrites the buffer
    }
}

// ------------------------------------------------------------------------------------------------
// Round 2: the same four operators with EIGHT consecutive samples per lane (float only).
//  * rows move as 32-byte vectors (ld.global.v8 / st.global.v8, sm_100: LDG.E.256 / STG.E.256): a warp
//    request still covers 1 KB of contiguous row;
//  * every window vector (16 bytes) read from shared memory now feeds 8 outputs instead of 4:
//    (K + 10) / 4 window loads per 8 outputs instead of 2 (K + 6) / 4 -- K = 28: 10 against 16 -- which
//    is what bounded the 4-sample kernel (70 % of the shared-memory wavefront peak at 0.71 of HBM);
//  * the shared copy of the row is skewed: one unused 16-byte slot after every 8 vectors, so that the
//    32-byte lane stride of these loads / stores touches every bank once per quarter warp (without the
//    skew they are 2-way conflicts, the variant round 1 measured and rejected).
// Rows must start on 32-byte boundaries (T % 8 == 0), T <= 256 NCH, K <= KMAX <= 32.
// ------------------------------------------------------------------------------------------------
struct alignas(32) Vec32f {
    float t[8];
};
__device__ __forceinline__ Vec32f ldg256(const float *p) {
    Vec32f v;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v.t[0]), "=f"(v.t[1]), "=f"(v.t[2]), "=f"(v.t[3]), "=f"(v.t[4]), "=f"(v.t[5]), "=f"(v.t[6]),
                   "=f"(v.t[7])
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void stg256(float *p, const Vec32f &v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v.t[0]), "f"(v.t[1]),
                 "f"(v.t[2]), "f"(v.t[3]), "f"(v.t[4]), "f"(v.t[5]), "f"(v.t[6]), "f"(v.t[7])
                 : "memory");
}

template <int KMAX, int NCH>
struct RowsConv8Layout {
    static constexpr int NPV = (KMAX - 1 + 3) / 4;                  // vectors the taps reach beyond an output pair
    static constexpr int FRAME = 8;                                 // zero vectors on each side (>= NPV)
    static constexpr int NV = 64 * NCH;                             // data vectors (16 bytes each)
    static constexpr int SLOTS = ((NV + 2 * FRAME) / 8) * 9;        // with the skew slots
    static constexpr size_t WARP_BYTES = (size_t)SLOTS * 16 + (size_t)KMAX * sizeof(float);
    static_assert(NPV <= FRAME, "taps reach beyond the zero frame");
    // slot of logical vector v (v = -FRAME .. NV + FRAME - 1)
    __host__ __device__ static constexpr int slot(int v) { return (v + FRAME) + (v + FRAME) / 8; }
};

template <int OP, int KMAX, int NCH>
__global__ void __launch_bounds__(256, 2)
rows_conv8_kernel(const float *__restrict__ h, int64_t h_stride, const float *__restrict__ x, float *out,
                  int64_t V, int T, int K) {
    using L = RowsConv8Layout<KMAX, NCH>;
    constexpr bool ADJ = OP == OP_CONV_ADJ || OP == OP_HRFINTEG_ADJ;
    constexpr int NPV = L::NPV;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    Vec16<float> *row = reinterpret_cast<Vec16<float> *>(smem + (size_t)warp * L::WARP_BYTES);
    float *hs = reinterpret_cast<float *>(row + L::SLOTS);
    for (int i = lane; i < L::SLOTS; i += 32)
        row[i] = Vec16<float>{{0.f, 0.f, 0.f, 0.f}};              // zero frames (and everything else once)
    const int n8 = T / 8;
    float taps[KMAX];
    bool have_taps = false;
    for (int64_t v = (int64_t)blockIdx.x * nw + warp; v < V; v += (int64_t)gridDim.x * nw) {
        Vec32f d[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {                             // the whole row in flight
            const int i = c * 32 + lane;
            if (i < n8) {
                d[c] = ldg256(x + v * T + 8 * (int64_t)i);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) d[c].t[e] = 0.f;
            }
        }
        if (h_stride != 0 || !have_taps) {
            __syncwarp();
            for (int a = lane; a < KMAX; a += 32) hs[a] = a < K ? h[v * h_stride + a] : 0.f;
        }
        if (OP == OP_HRFINTEG) {                                    // cumsum first (pybold/linear.py:86)
            float carry = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
#pragma unroll
                for (int e = 1; e < 8; ++e) d[c].t[e] += d[c].t[e - 1];
                float inc = d[c].t[7];
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const float t = __shfl_up_sync(PB_FULL, inc, sft);
                    if (lane >= sft) inc += t;
                }
                float ex = __shfl_up_sync(PB_FULL, inc, 1);
                ex = (lane == 0 ? 0.f : ex) + carry;
#pragma unroll
                for (int e = 0; e < 8; ++e) d[c].t[e] += ex;
                carry += __shfl_sync(PB_FULL, inc, 31);
            }
            // the cumsum runs on into the zero fill beyond T: those inputs of the convolution must stay zero
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                if (c * 32 + lane >= n8) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) d[c].t[e] = 0.f;
                }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int v0 = 2 * (c * 32 + lane);
            row[L::slot(v0)] = Vec16<float>{{d[c].t[0], d[c].t[1], d[c].t[2], d[c].t[3]}};
            row[L::slot(v0 + 1)] = Vec16<float>{{d[c].t[4], d[c].t[5], d[c].t[6], d[c].t[7]}};
        }
        __syncwarp();
        if (h_stride != 0 || !have_taps) {
#pragma unroll
            for (int a = 0; a < KMAX; ++a) taps[a] = hs[a];
            have_taps = true;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
            if (c * 256 < T) {                                      // warp-uniform
                const int v0 = 2 * (c * 32 + lane);
#pragma unroll
                for (int w = 0; w < NPV + 2; ++w) {
                    // causal: vectors v0 - NPV .. v0 + 1; anti-causal: v0 .. v0 + 1 + NPV
                    const int vv = ADJ ? v0 + w : v0 - NPV + w;
                    const Vec16<float> xv = row[L::slot(vv)];
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            // sample index relative to the lane's first output: s = 4 (w - NPV) + t (causal), 4 w + t (adj)
                            const int sidx = ADJ ? 4 * w + t : 4 * (w - NPV) + t;
                            const int j = ADJ ? sidx - e : e - sidx;
                            if (j >= 0 && j < KMAX) acc[e] = fmaf(taps[j >= 0 && j < KMAX ? j : 0], xv.t[t], acc[e]);
                        }
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) d[c].t[e] = acc[e];
        }
        if (OP == OP_HRFINTEG_ADJ) {                                // reversed cumsum (pybold/linear.py:113)
            float carry = 0.f;
#pragma unroll
            for (int c = NCH - 1; c >= 0; --c) {
#pragma unroll
                for (int e = 6; e >= 0; --e) d[c].t[e] += d[c].t[e + 1];
                float inc = d[c].t[0];
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const float t = __shfl_down_sync(PB_FULL, inc, sft);
                    if (lane + sft < 32) inc += t;
                }
                float ex = __shfl_down_sync(PB_FULL, inc, 1);
                ex = (lane == 31 ? 0.f : ex) + carry;
#pragma unroll
                for (int e = 0; e < 8; ++e) d[c].t[e] += ex;
                carry += __shfl_sync(PB_FULL, inc, 0);
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i = c * 32 + lane;
            if (i < n8) stg256(out + v * T + 8 * (int64_t)i, d[c]);
        }
        __syncwarp();                                               // the next row overwrites the buffer
    }
}

}  // namespace pb
