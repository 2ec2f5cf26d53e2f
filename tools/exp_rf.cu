// Developer microbenchmark: register-file read cost of FFMA on sm_100a.
//  MODE 0: tap-stationary, 16 FFMA per tap (tap in the reuse cache, 2 register reads per FFMA)
//  MODE 1: tap changes every FFMA (3 register reads per FFMA)
//  MODE 2: like 0 with one FSEL (ALU pipe) after every 4 FFMA
//  MODE 3: like 1 with taps alternating between two registers
// Check the SASS (cuobjdump -sass) for the .reuse flags and register parities before trusting a number.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(float *sink, const float *in, int iters, long long *cyc) {
    float acc[16], d[16], h[16], acc2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = in[i * 128 + threadIdx.x]; acc2[i] = acc[i]; d[i] = in[(16 + i) * 128 + threadIdx.x]; h[i] = in[(32 + i) * 128 + threadIdx.x]; }
    float sel = in[0];
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (MODE == 0) acc[r] = fmaf(h[j], d[(r + j) & 15], acc[r]);
                if (MODE == 1) acc[r] = fmaf(h[(j + r) & 15], d[(r + 2 * j) & 15], acc[r]);
                if (MODE == 2) { acc[r] = fmaf(h[j], d[(r + j) & 15], acc[r]); if ((r & 3) == 3) sel = sel > acc[r] ? sel : d[r]; }
                if (MODE == 4) acc[r] = fmaf(h[j], d[(r + 2 * j) & 15], acc[r]);
                if (MODE == 5) acc[r] = fmaf(h[j], d[(r + 2 * j + 1) & 15], acc[r]);
                if (MODE == 6) acc[r] = fmaf(acc[r], h[0], h[1]);
                if (MODE == 7) acc[r] = fmaf(h[j], acc[(r + 1) & 15], acc[r]);
                if (MODE == 8) acc[r] = fmaf(h[j], d[r], acc[r]);
                if (MODE == 9) { if (j & 1) acc[r] = fmaf(h[j], d[(r + j) & 15], acc2[r]); else acc2[r] = fmaf(h[j], d[(r + j) & 15], acc[r]); }
                if (MODE == 3) acc[r] = fmaf(h[(r & 1) + 2 * (j & 7)], d[(r + j) & 15], acc[r]);
            }
        }
    }
    const long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0 && cyc) *cyc = c1 - c0;
    float s = sel;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *cyc; cudaMalloc(&cyc, 8); float *sink, *in; cudaMalloc(&sink, sms * 8 * 128 * 4); cudaMalloc(&in, 48 * 128 * 4); cudaMemset(in, 0, 48 * 128 * 4);
    for (int bps : {1, 2, 3, 4, 8}) {
        const int blocks = sms * bps, iters = 4000;
        k<MODE><<<blocks, 128>>>(sink, in, 200000, nullptr); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(sink, in, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fl = (double)blocks * 128 * iters * 256.0 * 2;
        long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s %d warps/SMSP: %7.2f ms  %6.2f Tflop/s  | %.3f cycles per FFMA per SMSP, %.0f MHz\n", name, bps, best, fl / best / 1e9,
               (double)hc / ((double)iters * 256 * bps), (double)hc / best / 1e3);
    }
    cudaFree(sink); cudaFree(in);
}

int main() {
    run<0>("tap-stationary (2 RF reads)");
    run<1>("tap changes every FFMA (3 RF reads)");
    run<2>("tap-stationary + FSEL every 4");
    run<3>("two alternating taps (3 RF reads)");
    run<6>("acc = acc*m + c (1 RF read)");
    run<7>("acc = tap*acc[r+1] + acc");
    run<8>("acc = tap*d[r] + acc");
    run<9>("ping-pong accumulators (dst != src)");
    run<4>("tap-stationary, data index r+2j");
    run<5>("tap-stationary, data index r+2j+1");
    return 0;
}
