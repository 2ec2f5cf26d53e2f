// Developer microbenchmark: FFMA rate of the solver's register tile (R accumulators, K taps,
// sliding data window) as ptxas compiles it, without shuffles / scans.
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int K, int MODE>
__global__ void __launch_bounds__(128, 3) tile(float *sink, int iters) {
    float d[R + K - 1], h[K], acc[R], acc2[R];
    const float t = threadIdx.x * 1e-6f;
#pragma unroll
    for (int i = 0; i < R + K - 1; ++i) d[i] = t + i * 0.01f;
#pragma unroll
    for (int j = 0; j < K; ++j) h[j] = 0.01f * (j + 1) + t;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = 0.f; acc2[r] = 0.f; }
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fmaf(h[j], d[r + K - 1 - j], acc[r]);
        } else {   // even taps into acc, odd taps into acc2
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (j & 1) acc2[r] = fmaf(h[j], d[r + K - 1 - j], acc2[r]);
                    else acc[r] = fmaf(h[j], d[r + K - 1 - j], acc[r]);
                }
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] += acc2[r];
        }
        // feed the results back into the window (keeps everything live, costs R + K - 1 FMUL-free moves)
#pragma unroll
        for (int i = 0; i < R + K - 1; ++i) d[i] = acc[i % R] * 1e-3f + d[i];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < R + K - 1; ++i) s += d[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int R, int K, int MODE>
void run(const char *name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 3, iters = 4000;
    float *sink; cudaMalloc(&sink, blocks * 128 * 4);
    tile<R, K, MODE><<<blocks, 128>>>(sink, 10); cudaDeviceSynchronize();
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, tile<R, K, MODE>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); tile<R, K, MODE><<<blocks, 128>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double fl = (double)blocks * 128 * iters * (double)R * K * 2;
    printf("%-36s regs=%3d spill=%3zu : %7.2f ms  %6.2f Tflop/s (conv FFMA only)\n", name, fa.numRegs, (size_t)fa.localSizeBytes, best, fl / best / 1e9);
    cudaFree(sink);
}

int main() {
    run<19, 19, 0>("R19 K19 single accumulators");
    run<19, 19, 1>("R19 K19 split even/odd accumulators");
    run<10, 19, 0>("R10 K19 single accumulators");
    run<10, 19, 1>("R10 K19 split even/odd accumulators");
    run<20, 27, 0>("R20 K27 single accumulators");
    run<20, 27, 1>("R20 K27 split even/odd accumulators");
    return 0;
}
