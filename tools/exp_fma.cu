// Developer microbenchmark: FP32 FMA-pipe ceilings on sm_100a (FFMA vs packed FFMA2, with and
// without competing ALU / SHFL instructions).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_fma tools/exp_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *sink, int iters) {
    const float t = threadIdx.x * 1e-9f;
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(t + i, t + i + 0.5f);
    const float2 m = make_float2(0.999f + blockIdx.x * 1e-9f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    int x = threadIdx.x, y2 = blockIdx.x;
    float s = t;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // 16 FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        } else if (MODE == 1) {   // 8 FFMA2 (= 16 FMA)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
        } else if (MODE == 2) {   // 8 FFMA2 + 8 integer ALU
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = __ffma2_rn(a[i], m, c); x = (x ^ y2) + (x >> 3); y2 = y2 * 3 + x; }
        } else if (MODE == 3) {   // 16 FFMA + 8 integer ALU
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); x = (x ^ y2) + (x >> 3); y2 = y2 * 3 + x; }
        } else if (MODE == 4) {   // 8 FFMA2 + 4 SHFL
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = __ffma2_rn(a[i], m, c); if (i & 1) s += __shfl_up_sync(0xffffffffu, s, 1); }
        } else if (MODE == 5) {   // 16 FFMA + 4 SHFL
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); if (i & 1) s += __shfl_up_sync(0xffffffffu, s, 1); }
        } else if (MODE == 6) {   // 8 FFMA2 with 3 distinct register operands (tap * data + acc), like the solver
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(m, a[(i + 1) & 7], a[i]);
        } else if (MODE == 7) {   // 16 FFMA, 3 distinct register operands
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(m.x, a[(i + 1) & 7].x, a[i].x); a[i].y = fmaf(m.y, a[(i + 1) & 7].y, a[i].y); }
        } else if (MODE == 8) {   // 16 FFMA, acc and data in adjacent registers (different banks)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(m.x, a[i].y, a[i].x); a[(i + 3) & 7].y = fmaf(m.y, a[(i + 3) & 7].x, a[(i + 3) & 7].y); }
        } else if (MODE == 9) {   // 16 FFMA, acc and data same parity
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(m.x, a[(i + 1) & 7].x, a[i].x); a[(i + 3) & 7].y = fmaf(m.y, a[(i + 4) & 7].y, a[(i + 3) & 7].y); }
        }
    }
    float r = s + x + y2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm, iters = 1 << 15;
    float *sink; cudaMalloc(&sink, blocks * 256 * 4);
    k<MODE><<<blocks, 256>>>(sink, 256);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(sink, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double fl = (double)blocks * 256 * 16 * 2 * iters;
    printf("%-44s %d blk/SM (%2d warps/SM): %7.2f ms  %6.2f Tflop/s\n", name, blocks_per_sm, blocks_per_sm * 8, best, fl / best / 1e9);
    cudaFree(sink);
}

int main() {
    for (int b : {4}) {
        run<0>("16 FFMA", b);
        run<1>("8 FFMA2", b);
        run<2>("8 FFMA2 + 16 int ALU", b);
        run<3>("16 FFMA + 16 int ALU", b);
        run<4>("8 FFMA2 + 4 SHFL + 4 FADD", b);
        run<5>("16 FFMA + 4 SHFL + 4 FADD", b);
        run<6>("8 FFMA2 tap*data+acc", b);
        run<7>("16 FFMA tap*data+acc", b);
        run<8>("16 FFMA tap*data+acc, adjacent regs", b);
        run<9>("16 FFMA tap*data+acc, same parity", b);
    }
    return 0;
}
