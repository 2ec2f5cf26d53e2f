// Developer experiment (not shipped): times fast_bd_kernel instantiations on random data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I pybold_b200/csrc \
//        -o tools/exp_bd tools/exp_bd.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pb_fast.cuh"

using namespace pb;

static HrfGrid make_grid(double t_r, double dur) {
    HrfGrid g; const int N = (int)(dur / 0.001); const int stride = (int)(t_r / 0.001);
    g.t_step = dur / (double)(N - 1); g.stride = stride; g.K = (N + stride - 1) / stride; return g;
}

template <int R, int KMAX, bool CIRC, int WARPS, int MINB>
void run(const char *name, int64_t V, int T, double t_r, int nb_iter) {
    BdArgs<float> a;
    a.grid = make_grid(t_r, 20.0);
    a.K = a.grid.K; a.T = T; a.V = V;
    std::vector<float> y((size_t)V * T);
    srand(1);
    for (auto &v : y) v = (float)rand() / RAND_MAX - 0.3f;
    float *dy, *lb, *th, *x, *z, *dz, *h, *theta, *J, *r, *g; int32_t *nt;
    cudaMalloc(&dy, V * T * 4); cudaMemcpy(dy, y.data(), V * T * 4, cudaMemcpyHostToDevice);
    float one = 1.7f, two = 2.0f;
    cudaMalloc(&lb, 4); cudaMemcpy(lb, &one, 4, cudaMemcpyHostToDevice);
    cudaMalloc(&th, 4); cudaMemcpy(th, &two, 4, cudaMemcpyHostToDevice);
    cudaMalloc(&x, V * T * 4); cudaMalloc(&z, V * T * 4); cudaMalloc(&dz, V * T * 4);
    cudaMalloc(&h, V * a.K * 4); cudaMalloc(&theta, V * 4);
    cudaMalloc(&J, V * (nb_iter + 2) * 4); cudaMalloc(&r, V * (nb_iter + 2) * 4); cudaMalloc(&g, V * (nb_iter + 2) * 4);
    cudaMalloc(&nt, V * 4);
    a.y = dy; a.lbda = lb; a.lbda_stride = 0; a.theta0 = th; a.theta0_stride = 0; a.z0 = nullptr;
    a.theta_lo = 0.6; a.theta_hi = 1.9; a.nb_iter = nb_iter; a.early_stopping = 0; a.wind = 4; a.tol = 1e-12;
    a.out_x = x; a.out_z = z; a.out_dz = dz; a.out_h = h; a.out_theta = theta; a.out_J = J; a.out_r = r; a.out_g = g;
    a.out_ntrace = nt;
    auto kern = fast_bd_kernel<float, R, KMAX, CIRC, WARPS, MINB>;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    int rc = fast_bd_launch<float, R, KMAX, CIRC, WARPS, MINB>(a, 0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        rc = fast_bd_launch<float, R, KMAX, CIRC, WARPS, MINB>(a, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const size_t beta_bytes = ((size_t)nb_iter * 4 + 15) & ~(size_t)15;
    const size_t smem = beta_bytes + (size_t)WARPS * pb_scratch_doubles(KMAX) * 8;
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    const int K = a.K; const double MAC = (double)T * K - K * (K - 1) / 2.0;
    const double flops = (double)V * ((double)(nb_iter + 1) * nb_iter * (4 * MAC + 11 * T));
    printf("%-28s rc=%d regs=%3d spill=%4zu occ=%d blk/SM (%2d warps) T=%d K=%d V=%lld: %8.2f ms %9.0f vox/s %6.2f Tflop/s\n",
           name, rc, fa.numRegs, (size_t)fa.localSizeBytes, occ, occ * WARPS, T, K, (long long)V, best, V / best * 1e3, flops / best / 1e9);
    cudaFree(dy); cudaFree(x); cudaFree(z); cudaFree(dz); cudaFree(h); cudaFree(theta); cudaFree(J); cudaFree(r); cudaFree(g); cudaFree(nt);
}

#define RUN(R, K, C, W, M, V, T, TR) run<R, K, C, W, M>("R" #R " K" #K " " #C " W" #W " M" #M, V, T, TR, 100)

int main() {
    RUN(8, 28, false, 4, 3, 40000, 240, 0.75);
    RUN(8, 28, false, 4, 4, 40000, 240, 0.75);
    RUN(8, 28, false, 4, 5, 40000, 240, 0.75);
    RUN(10, 20, false, 4, 4, 40000, 310, 1.0);
    RUN(10, 20, false, 4, 5, 40000, 310, 1.0);
    RUN(10, 32, false, 4, 3, 40000, 300, 0.72);
    RUN(10, 32, false, 4, 4, 40000, 300, 0.72);
    RUN(10, 32, false, 4, 5, 40000, 300, 0.72);
    RUN(20, 20, false, 4, 2, 16000, 610, 1.0);
    RUN(20, 20, false, 4, 3, 16000, 610, 1.0);
    RUN(20, 32, false, 4, 2, 16000, 610, 0.72);
    RUN(20, 32, false, 4, 3, 16000, 610, 0.72);
    RUN(40, 32, false, 4, 1, 8000, 1210, 0.72);
    RUN(40, 32, false, 4, 2, 8000, 1210, 0.72);
    RUN(40, 28, true, 6, 1, 8000, 1200, 0.72);
    RUN(40, 28, true, 8, 1, 8000, 1200, 0.72);
    return 0;
}
