// The unrolled R x K register tile of the solvers: acc[r] += h[j] * x[r -+ j] for all taps j and
// the R samples a lane owns (x = own samples `a`, or the halo beyond them).
//
// The SOURCE ORDER of the R x K FFMA matters: ptxas keeps part of it, and the order decides how many
// FFMA find an operand in the operand-reuse cache.  The kernels are bound by register-file read
// bandwidth (one 32-bit operand vector per bank and cycle, profiles/r01_rf_bandwidth.txt), so an
// FFMA with three register reads costs 1.5 issue cycles and one with two reads costs one.  The
// orders below (tap direction, sample direction, accumulator block size RB, or data-stationary DS)
// are picked per kernel family with that model on the SASS and then timed on a B200.
// Changing the order changes the FP32 summation order of a sample, nothing else.
#pragma once

namespace pb {

// Predicated FMA: acc += h * v only where ok != 0 (one predicate per halo hop, constant over the kernel).
// The lanes at the ends of a voxel's lane group have no neighbour to take a halo from; instead of zeroing
// every halo value with a select after the shuffle, the FFMA that would consume it is predicated off.
__device__ __forceinline__ float fma_if(float h, float v, float acc, unsigned ok) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p fma.rn.f32 %0, %1, %2, %0;\n\t}"
        : "+f"(acc) : "f"(h), "f"(v), "r"(ok));
    return acc;
}
__device__ __forceinline__ double fma_if(double h, double v, double acc, unsigned ok) {
    return ok ? fma(h, v, acc) : acc;
}

struct TileOrder {
    int jdesc;   // taps from K-1 down to JS instead of JS .. K-1
    int rdesc;   // samples from R-1 down to 0
    int rb;      // accumulator block: all taps for RB samples, then the next RB (0 = all R at once)
    int ds;      // 1 / 2: data-stationary (all taps of one window position), ascending / descending
};

// causal: acc[r] += h[j] * x[r - j],  x[i] = a[i] (i >= 0) or halo[-i - 1]
// PRED: halo values are raw (not zeroed at the ends of the lane group); ok[d - 1] != 0 where the lane d
// positions away exists
template <typename real, int R, int KMAX, int NHALO, int JS, int JDESC, int RDESC, int RB_, int DS, bool PRED = false>
__device__ __forceinline__ void tile_conv(const real (&h)[KMAX], const real (&a)[R],
                                          const real (&halo)[NHALO], real (&acc)[R],
                                          const unsigned *ok = nullptr) {
    // PRED: the first FFMA of an accumulator must not be a predicated one (its initial value would need a
    // select): the tap-JS terms on own samples go first
    if constexpr (PRED) {
#pragma unroll
        for (int r = JS; r < R; ++r) acc[r] = fma(h[JS], a[r - JS], acc[r]);
    }
    if constexpr (DS != 0) {
#pragma unroll
        for (int ii = -(KMAX - 1); ii < R; ++ii) {
            const int i = DS == 2 ? (R - 1) - (ii + KMAX - 1) : ii;
            const real val = i >= 0 ? a[i >= 0 ? i : 0] : halo[i >= 0 ? 0 : -i - 1];
#pragma unroll
            for (int jj = JS; jj < KMAX; ++jj) {
                const int j = JDESC ? KMAX - 1 + JS - jj : jj;
                const int r = i + j;
                if (PRED && j == JS && i >= 0) continue;
                if (r >= 0 && r < R) {
                    if (PRED && i < 0)
                        acc[r >= 0 && r < R ? r : 0] = fma_if(h[j], val, acc[r >= 0 && r < R ? r : 0], ok[i < 0 ? (-i + R - 1) / R - 1 : 0]);
                    else
                        acc[r >= 0 && r < R ? r : 0] = fma(h[j], val, acc[r >= 0 && r < R ? r : 0]);
                }
            }
        }
    } else {
        constexpr int RB = RB_ > 0 ? RB_ : R;
#pragma unroll
        for (int rb = 0; rb < R; rb += RB) {
#pragma unroll
            for (int jj = JS; jj < KMAX; ++jj) {
                const int j = JDESC ? KMAX - 1 + JS - jj : jj;
#pragma unroll
                for (int rr = rb; rr < (rb + RB < R ? rb + RB : R); ++rr) {
                    const int r = RDESC ? R - 1 - rr : rr;
                    const int idx = r - j;
                    const real val = idx >= 0 ? a[idx >= 0 ? idx : 0] : halo[idx >= 0 ? 0 : -idx - 1];
                    if (PRED && j == JS && idx >= 0) continue;
                    if (PRED && idx < 0) acc[r] = fma_if(h[j], val, acc[r], ok[idx < 0 ? (-idx + R - 1) / R - 1 : 0]);
                    else acc[r] = fma(h[j], val, acc[r]);
                }
            }
        }
    }
}

// anti-causal: acc[r] += h[j] * x[r + j],  x[i] = a[i] (i < R) or halo[i - R]
template <typename real, int R, int KMAX, int NHALO, int JS, int JDESC, int RDESC, int RB_, int DS, bool PRED = false>
__device__ __forceinline__ void tile_corr(const real (&h)[KMAX], const real (&a)[R],
                                          const real (&halo)[NHALO], real (&acc)[R],
                                          const unsigned *ok = nullptr) {
    if constexpr (PRED) {
#pragma unroll
        for (int r = 0; r + JS < R; ++r) acc[r] = fma(h[JS], a[r + JS], acc[r]);
    }
    if constexpr (DS != 0) {
#pragma unroll
        for (int ii = 0; ii < R + KMAX - 1; ++ii) {
            const int i = DS == 2 ? (R + KMAX - 2) - ii : ii;
            const real val = i < R ? a[i < R ? i : 0] : halo[i < R ? 0 : i - R];
#pragma unroll
            for (int jj = JS; jj < KMAX; ++jj) {
                const int j = JDESC ? KMAX - 1 + JS - jj : jj;
                const int r = i - j;
                if (PRED && j == JS && i < R) continue;
                if (r >= 0 && r < R) {
                    if (PRED && i >= R)
                        acc[r >= 0 && r < R ? r : 0] = fma_if(h[j], val, acc[r >= 0 && r < R ? r : 0], ok[i >= R ? i / R - 1 : 0]);
                    else
                        acc[r >= 0 && r < R ? r : 0] = fma(h[j], val, acc[r >= 0 && r < R ? r : 0]);
                }
            }
        }
    } else {
        constexpr int RB = RB_ > 0 ? RB_ : R;
#pragma unroll
        for (int rb = 0; rb < R; rb += RB) {
#pragma unroll
            for (int jj = JS; jj < KMAX; ++jj) {
                const int j = JDESC ? KMAX - 1 + JS - jj : jj;
#pragma unroll
                for (int rr = rb; rr < (rb + RB < R ? rb + RB : R); ++rr) {
                    const int r = RDESC ? R - 1 - rr : rr;
                    const int idx = r + j;
                    const real val = idx < R ? a[idx < R ? idx : 0] : halo[idx < R ? 0 : idx - R];
                    if (PRED && j == JS && idx < R) continue;
                    if (PRED && idx >= R) acc[r] = fma_if(h[j], val, acc[r], ok[idx >= R ? idx / R - 1 : 0]);
                    else acc[r] = fma(h[j], val, acc[r]);
                }
            }
        }
    }
}

}  // namespace pb
