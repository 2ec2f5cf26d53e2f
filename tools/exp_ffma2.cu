#include <cstdio>
#include <cuda_runtime.h>
// FFMA2 probe: packed fma.rn.f32x2 throughput vs scalar FFMA, with and without interleaved non-FMA work
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float s) {
    const int NC = 8;
    if (MODE == 0) {           // scalar FFMA: 16 chains
        float a[2 * NC];
        for (int i = 0; i < 2 * NC; ++i) a[i] = threadIdx.x * 1e-3f + i;
        float b = s, c = s * 0.5f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 2 * NC; ++i) a[i] = fmaf(a[i], b, c);
        }
        float r = 0;
        for (int i = 0; i < 2 * NC; ++i) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 1) {    // packed: 8 chains of pairs (same flops per loop trip)
        unsigned long long a[NC];
        for (int i = 0; i < NC; ++i) a[i] = ((unsigned long long)__float_as_uint(threadIdx.x * 1e-3f + i) << 32) | __float_as_uint(1.0f + i);
        unsigned long long b = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
        unsigned long long c = ((unsigned long long)__float_as_uint(s * 0.5f) << 32) | __float_as_uint(s * 0.25f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < NC; ++i) a[i] = fma2(a[i], b, c);
        }
        unsigned long long r = 0;
        for (int i = 0; i < NC; ++i) r ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(r ^ (r >> 32)));
    } else if (MODE == 2) {    // scalar FFMA + 1 shuffle per 4 FFMA (mix like the solver)
        float a[2 * NC];
        for (int i = 0; i < 2 * NC; ++i) a[i] = threadIdx.x * 1e-3f + i;
        float b = s, c = s * 0.5f, sh = s;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < 2 * NC; ++i) a[i] = fmaf(a[i], b, c);
#pragma unroll
                for (int q = 0; q < 4; ++q) sh = __shfl_up_sync(0xffffffffu, sh, 1) + 0.0f * q;
            }
        }
        float r = sh;
        for (int i = 0; i < 2 * NC; ++i) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {                   // packed + the same shuffles
        unsigned long long a[NC];
        for (int i = 0; i < NC; ++i) a[i] = ((unsigned long long)__float_as_uint(threadIdx.x * 1e-3f + i) << 32) | __float_as_uint(1.0f + i);
        unsigned long long b = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
        unsigned long long c = ((unsigned long long)__float_as_uint(s * 0.5f) << 32) | __float_as_uint(s * 0.25f);
        float sh = s;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < NC; ++i) a[i] = fma2(a[i], b, c);
#pragma unroll
                for (int q = 0; q < 4; ++q) sh = __shfl_up_sync(0xffffffffu, sh, 1) + 0.0f * q;
            }
        }
        unsigned long long r = 0;
        for (int i = 0; i < NC; ++i) r ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(r ^ (r >> 32))) + sh;
    }
}
template <int MODE>
void run(const char *name, float *out, int sms) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<sms * 4, 256>>>(out, iters, 1.0000001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double flops = (double)sms * 4 * 256 * iters * 8 * 16 * 2;
    printf("%-40s %8.3f ms  %7.2f Tflop/s  (%s)\n", name, best, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, sms * 4 * 256 * 4);
    run<0>("scalar FFMA, 16 chains", out, sms);
    run<1>("packed fma.rn.f32x2, 8 pair chains", out, sms);
    run<2>("scalar FFMA + 4 SHFL per 16 FFMA", out, sms);
    run<3>("packed f32x2 + 4 SHFL per 8 FFMA2", out, sms);
    return 0;
}
