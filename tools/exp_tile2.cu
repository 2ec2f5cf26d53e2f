// Developer microbenchmark: FFMA rate of the bare R x K register tile as ptxas schedules it, for
// several source-level orderings (see profiles/r01_rf_bandwidth.txt).  cycles are clock64() of one warp.
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int K, int MODE>
__global__ void __launch_bounds__(128, 3) tile(float *sink, const float *__restrict__ hin, int iters, int zmask, long long *cyc) {
    float d[R + K - 1], h[K], acc[R];
    const float t = threadIdx.x * 1e-6f;
#pragma unroll
    for (int i = 0; i < R + K - 1; ++i) d[i] = t + i * 0.01f;
#pragma unroll
    for (int j = 0; j < K; ++j) h[j] = hin[j * 32 + threadIdx.x];
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 4) { if (zmask) asm volatile("exit;"); }
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        if (MODE == 0 || MODE == 4) {
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fmaf(h[j], d[r + K - 1 - j], acc[r]);
        } else if (MODE == 1) {       // tap j+1 depends on an accumulator of tap j (LOP3 with an opaque zero)
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float tap = j ? __int_as_float(__float_as_int(h[j]) | (__float_as_int(acc[R - 1]) & zmask)) : h[j];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fmaf(tap, d[r + K - 1 - j], acc[r]);
            }
        } else if (MODE == 2) {       // accumulator-major source order
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < K; ++j) acc[r] = fmaf(h[j], d[r + K - 1 - j], acc[r]);
        }
#pragma unroll
        for (int i = 0; i < R + K - 1; ++i) d[i] = acc[i % R] * 1e-3f + d[i];
    }
    const long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = c1 - c0;
    float s = 0;
#pragma unroll
    for (int i = 0; i < R + K - 1; ++i) s += d[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int R, int K, int MODE>
void run(const char *name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *sink, *hin; long long *cyc, hc;
    cudaMalloc(&sink, sms * 3 * 128 * 4); cudaMalloc(&hin, 64 * 32 * 4 * 4); cudaMemset(hin, 0, 64 * 32 * 4 * 4); cudaMalloc(&cyc, 8);
    for (int bps : {1, 3}) {
        const int iters = 20000;
        tile<R, K, MODE><<<sms * bps, 128>>>(sink, hin, iters, 0, cyc);
        tile<R, K, MODE><<<sms * bps, 128>>>(sink, hin, iters, 0, cyc);
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s %d warps/SMSP: %.3f cycles per tile FFMA per SMSP\n", name, bps, (double)hc / ((double)iters * R * K * bps));
    }
}

int main() {
    run<19, 19, 0>("R19 K19 tap-major source (ptxas default)");
    run<19, 19, 1>("R19 K19 tap chain (LOP3)");
    run<19, 19, 2>("R19 K19 accumulator-major source");
    run<19, 19, 4>("R19 K19 tap-major + loop-top exit");
    return 0;
}
