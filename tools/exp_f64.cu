// Developer microbenchmark: latency / throughput of DFMA and SHFL on sm_100a (1 warp per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void dfma(double *sink, double m, double c, int iters, long long *cyc) {
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    const long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = c1 - c0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void shfl_chain(float *sink, int iters, long long *cyc) {
    float v = threadIdx.x;
    const long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u) v += __shfl_up_sync(0xffffffffu, v, 1, 16);
    const long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = c1 - c0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
__global__ void fadd_chain(float *sink, float c, int iters, long long *cyc) {
    float v = threadIdx.x;
    const long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 16; ++u) v += c;
    const long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = c1 - c0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink; float *fs; long long *cyc, hc; cudaMalloc(&sink, sms * 1024 * 8); cudaMalloc(&fs, sms * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int threads : {128, 512}) {
        dfma<1><<<sms, threads>>>(sink, 0.999, 1e-3, iters, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA 1 dependent chain, %d warps/SMSP: %.1f cycles per DFMA\n", threads / 128, (double)hc / (iters * 16.0));
        dfma<8><<<sms, threads>>>(sink, 0.999, 1e-3, iters, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA 8 independent chains, %d warps/SMSP: %.2f cycles per DFMA per warp\n", threads / 128, (double)hc / (iters * 16.0 * 8));
    }
    shfl_chain<<<sms, 128>>>(fs, iters, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("SHFL.UP + FADD dependent chain: %.1f cycles per step\n", (double)hc / (iters * 16.0));
    fadd_chain<<<sms, 128>>>(fs, 1.5f, iters, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FADD dependent chain: %.1f cycles per FADD\n", (double)hc / (iters * 16.0));
    return 0;
}
