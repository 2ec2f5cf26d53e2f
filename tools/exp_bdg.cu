// Developer experiment (not shipped): times fast_bdg_kernel instantiations and checks them against
// fast_bd_kernel (G = 32) on the same random data.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "pb_fast.cuh"
#include "pb_fastg.cuh"
#include "pb_fastc.cuh"

using namespace pb;

static HrfGrid make_grid(double t_r, double dur) {
    HrfGrid g; const int N = (int)(dur / 0.001); const int stride = (int)(t_r / 0.001);
    g.t_step = dur / (double)(N - 1); g.stride = stride; g.K = (N + stride - 1) / stride; return g;
}

struct Bufs {
    BdArgs<float> a; std::vector<float> z, theta; int64_t V; int T;
    float *dy, *lb, *th, *x, *dz_, *zz, *h, *theta_d, *J, *r, *g; int32_t *nt;
    void alloc(int64_t V_, int T_, double t_r, int nb_iter) {
        V = V_; T = T_;
        a.grid = make_grid(t_r, 20.0); a.K = a.grid.K; a.T = T; a.V = V;
        std::vector<float> y((size_t)V * T);
        srand(1);
        // block-like signal + noise so that the theta step has something to fit
        for (int64_t v = 0; v < V; ++v) {
            int on = rand() % (T - 40);
            for (int i = 0; i < T; ++i) {
                float s = (i > on + 4 && i < on + 30) ? 1.0f : 0.0f;
                y[v * T + i] = s + 0.3f * ((float)rand() / RAND_MAX - 0.5f);
            }
        }
        {   // real benchmark voxels when the file is there (tiled)
            char path[64]; snprintf(path, sizeof path, "tools/y%d.bin", T);
            FILE *f = fopen(path, "rb");
            if (f) {
                std::vector<float> u; fseek(f, 0, SEEK_END); long n = ftell(f) / 4; fseek(f, 0, SEEK_SET);
                u.resize(n); if (fread(u.data(), 4, n, f) != (size_t)n) n = 0; fclose(f);
                const long nv = n / T;
                if (nv > 0) { for (int64_t v = 0; v < V; ++v) for (int i = 0; i < T; ++i) y[v * T + i] = u[(v % nv) * T + i]; printf("using %s (%ld voxels)\n", path, nv); }
            }
        }
        cudaMalloc(&dy, V * T * 4); cudaMemcpy(dy, y.data(), V * T * 4, cudaMemcpyHostToDevice);
        float one = 1.7f, two = 2.0f;
        cudaMalloc(&lb, 4); cudaMemcpy(lb, &one, 4, cudaMemcpyHostToDevice);
        cudaMalloc(&th, 4); cudaMemcpy(th, &two, 4, cudaMemcpyHostToDevice);
        cudaMalloc(&x, V * T * 4); cudaMalloc(&zz, V * T * 4); cudaMalloc(&dz_, V * T * 4);
        cudaMalloc(&h, V * a.K * 4); cudaMalloc(&theta_d, V * 4);
        cudaMalloc(&J, V * (nb_iter + 2) * 4); cudaMalloc(&r, V * (nb_iter + 2) * 4); cudaMalloc(&g, V * (nb_iter + 2) * 4);
        cudaMalloc(&nt, V * 4);
        a.y = dy; a.lbda = lb; a.lbda_stride = 0; a.theta0 = th; a.theta0_stride = 0; a.z0 = nullptr;
        a.theta_lo = 0.6; a.theta_hi = 1.9; a.nb_iter = nb_iter; a.early_stopping = 0; a.wind = 4; a.tol = 1e-12;
        a.out_x = x; a.out_z = zz; a.out_dz = dz_; a.out_h = h; a.out_theta = theta_d; a.out_J = J; a.out_r = r; a.out_g = g;
        a.out_ntrace = nt;
    }
    void fetch() {
        z.resize((size_t)V * T); theta.resize(V);
        cudaMemcpy(z.data(), zz, V * T * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(theta.data(), theta_d, V * 4, cudaMemcpyDeviceToHost);
    }
    void free_all() { cudaFree(dy); cudaFree(x); cudaFree(zz); cudaFree(dz_); cudaFree(h); cudaFree(theta_d); cudaFree(J); cudaFree(r); cudaFree(g); cudaFree(nt); }
};

static std::vector<float> ref_z, ref_theta;

template <typename F, typename Kern>
void time_it(const char *name, Bufs &b, F launch, Kern kern, int warps, size_t smem, int nb_iter, bool is_ref) {
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    cudaMemset(b.zz, 0, b.V * b.T * 4);
    int rc = launch(); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); rc = launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    b.fetch();
#ifdef PB_DEBUG_EVALS
    { std::vector<int32_t> ne(b.V); cudaMemcpy(ne.data(), b.nt, b.V * 4, cudaMemcpyDeviceToHost);
      double s = 0; int mx = 0; for (auto e : ne) { s += e; if (e > mx) mx = e; }
      printf("    theta evals per voxel: mean %.1f max %d (per outer iteration %.2f)\n", s / b.V, mx, s / b.V / nb_iter); }
#endif
    double dz = 0, dt = 0, nz = 0;
    if (is_ref || ref_z.size() != b.z.size()) { ref_z = b.z; ref_theta = b.theta; }
    for (size_t i = 0; i < b.z.size(); ++i) { dz = fmax(dz, fabs(b.z[i] - ref_z[i])); nz = fmax(nz, fabs(ref_z[i])); }
    for (size_t i = 0; i < b.theta.size(); ++i) dt = fmax(dt, fabs(b.theta[i] - ref_theta[i]));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem);
    const int K = b.a.K, T = b.T; const double MAC = (double)T * K - K * (K - 1) / 2.0;
    const double flops = (double)b.V * ((double)(nb_iter + 1) * nb_iter * (4 * MAC + 11 * T));
    printf("%-30s rc=%d regs=%3d spill=%4zu occ=%d (%2d warps) T=%d K=%d: %8.2f ms %9.0f vox/s %6.2f Tflop/s | z err %.2e (max %.2f) theta err %.2e\n",
           name, rc, fa.numRegs, (size_t)fa.localSizeBytes, occ, occ * warps, T, K, best, b.V / best * 1e3, flops / best / 1e9, dz, nz, dt);
}

template <int R, int KMAX, bool CIRC, int WARPS, int MINB>
void run32(const char *name, Bufs &b, int nb_iter, bool is_ref) {
    auto kern = fast_bd_kernel<float, R, KMAX, CIRC, WARPS, MINB>;
    const size_t smem = (((size_t)nb_iter * 4 + 15) & ~(size_t)15) + (size_t)WARPS * pb_scratch_doubles(KMAX) * 8;
    time_it(name, b, [&] { return fast_bd_launch<float, R, KMAX, CIRC, WARPS, MINB>(b.a, 0); }, kern, WARPS, smem, nb_iter, is_ref);
}
template <int R, int KMAX, int G, int TAIL, int WARPS, int MINB, int LEAN = 0, int SMH = 0>
void rung(const char *name, Bufs &b, int nb_iter, bool is_ref = false) {
    auto kern = fast_bdg_kernel<float, R, KMAX, G, TAIL, WARPS, MINB, LEAN, SMH>;
    const size_t smem = (((size_t)nb_iter * 4 + 15) & ~(size_t)15) + (size_t)WARPS * fastg_warp_bytes<float, R, KMAX, G, LEAN, SMH>();
    time_it(name, b, [&] { return fast_bdg_launch<float, R, KMAX, G, TAIL, WARPS, MINB, LEAN, SMH>(b.a, 0); }, kern, WARPS, smem, nb_iter, is_ref);
}

template <int R, int KMAX, int NW, int MINB>
void runc(const char *name, Bufs &b, int nb_iter) {
    auto kern = fast_bdc_kernel<float, R, KMAX, NW, MINB>;
    const size_t smem = CtaLayout<float, R, KMAX, NW>::bytes(nb_iter);
    time_it(name, b, [&] { return fast_bdc_launch<float, R, KMAX, NW, MINB>(b.a, 0); }, kern, NW, smem, nb_iter, false);
}

int main(int argc, char **argv) {
    const int set = argc > 1 ? atoi(argv[1]) : 0;
    if (set == 0) {
        Bufs b; b.alloc(42624, 300, 1.0, 100);
        rung<19, 20, 16, 8, 4, 3>("G16 R19 K20 T8 W4 M3", b, 100, true);
        b.free_all();
    } else if (set == 9) {      // round 2: CTA shapes under the work queue -- 13 one-warp CTAs (152 registers), 6 x 2, 12 x 1
        Bufs b; b.alloc(42624, 300, 1.0, 100);
        rung<19, 20, 16, 8, 4, 3>("G16 R19 W4 M3 (default)", b, 100, true);
        rung<19, 20, 16, 8, 1, 13>("G16 R19 W1 M13 (13 warps)", b, 100);
        rung<19, 20, 16, 8, 1, 12>("G16 R19 W1 M12", b, 100);
        rung<19, 20, 16, 8, 2, 6>("G16 R19 W2 M6", b, 100);
        rung<19, 20, 16, 8, 1, 14>("G16 R19 W1 M14 (14 warps)", b, 100);
        b.free_all();
    } else if (set == 10) {     // round 2: one-warp CTAs (W1 M12) against W4 M3 on other group variants
        {
            Bufs b; b.alloc(21312, 240, 0.75, 100);
            rung<15, 28, 16, 8, 4, 3>("G16 R15 K28 W4 M3 (T=240)", b, 100, true);
            rung<15, 28, 16, 8, 1, 12>("G16 R15 K28 W1 M12 (T=240)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(21312, 200, 1.0, 100);
            rung<13, 20, 16, 13, 4, 3>("G16 R13 K20 W4 M3 (T=200)", b, 100, true);
            rung<13, 20, 16, 13, 1, 12>("G16 R13 K20 W1 M12 (T=200)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(21312, 320, 1.0, 100);
            rung<20, 20, 16, 20, 4, 3>("G16 R20 K20 W4 M3 (T=320)", b, 100, true);
            rung<20, 20, 16, 20, 1, 12>("G16 R20 K20 W1 M12 (T=320)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 128, 1.0, 100);
            rung<16, 20, 8, 16, 4, 3>("G8 R16 K20 W4 M3 (T=128)", b, 100, true);
            rung<16, 20, 8, 16, 1, 12>("G8 R16 K20 W1 M12 (T=128)", b, 100);
            b.free_all();
        }
    } else if (set == 11) {     // round 2: one-warp CTAs for the four-voxels-per-warp variants (G = 8, K <= 20)
        {
            Bufs b; b.alloc(28416, 64, 1.0, 100);
            rung<10, 20, 8, 10, 4, 3>("G8 R10 W4 M3 (T=64)", b, 100, true);
            rung<10, 20, 8, 10, 1, 12>("G8 R10 W1 M12 (T=64)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 100, 1.0, 100);
            rung<13, 20, 8, 13, 4, 3>("G8 R13 W4 M3 (T=100)", b, 100, true);
            rung<13, 20, 8, 13, 1, 12>("G8 R13 W1 M12 (T=100)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 150, 1.0, 100);
            rung<20, 20, 8, 20, 4, 3>("G8 R20 W4 M3 (T=150)", b, 100, true);
            rung<20, 20, 8, 20, 1, 12>("G8 R20 W1 M12 (T=150)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 190, 1.0, 100);
            rung<24, 20, 8, 24, 4, 3>("G8 R24 W4 M3 (T=190)", b, 100, true);
            rung<24, 20, 8, 24, 1, 12>("G8 R24 W1 M12 (T=190)", b, 100);
            b.free_all();
        }
    } else if (set == 4) {      // round 2: blocked shared-memory halo exchange, occupancy
        Bufs b; b.alloc(42624, 300, 1.0, 100);
        rung<19, 20, 16, 8, 4, 3>("G16 R19 K20 T8 W4 M3 (r01)", b, 100, true);
        rung<19, 20, 16, 8, 4, 3, 0, 2>("G16 R19 SMH2 M3", b, 100);
        rung<19, 20, 16, 8, 4, 4, 0, 2>("G16 R19 SMH2 M4", b, 100);
        rung<19, 20, 16, 8, 4, 4, 2, 2>("G16 R19 SMH2 LEAN2 M4", b, 100);
        rung<19, 20, 16, 8, 4, 3, 2, 2>("G16 R19 SMH2 LEAN2 M3", b, 100);
        rung<20, 20, 16, 8, 4, 3, 0, 2>("G16 R20 SMH2 M3", b, 100);
        rung<20, 20, 16, 8, 4, 4, 1, 2>("G16 R20 SMH2 LEAN1 M4", b, 100);
        b.free_all();
    } else if (set == 8) {      // round 2: short series with 28 taps -- four voxels per warp (G = 8) or two (G = 16)?
        {
            Bufs b; b.alloc(28416, 128, 0.72, 100);
            rung<16, 28, 8, 16, 4, 2>("G8  R16 K28 M2 (T=128)", b, 100, true);
            rung<8, 28, 16, 8, 4, 3>("G16 R8  K28 M3 (T=128)", b, 100);
            rung<8, 28, 16, 8, 4, 4>("G16 R8  K28 M4 (T=128)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 96, 0.72, 100);
            rung<13, 28, 8, 13, 4, 2>("G8  R13 K28 M2 (T=96)", b, 100, true);
            rung<6, 28, 16, 6, 4, 4>("G16 R6  K28 M4 (T=96)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 190, 0.72, 100);
            rung<24, 28, 8, 24, 4, 2>("G8  R24 K28 M2 (T=190)", b, 100, true);
            rung<12, 28, 16, 12, 4, 3>("G16 R12 K28 M3 (T=190)", b, 100);
            rung<12, 28, 16, 12, 4, 4>("G16 R12 K28 M4 (T=190)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 128, 1.0, 100);
            rung<16, 20, 8, 16, 4, 3>("G8  R16 K20 M3 (T=128)", b, 100, true);
            rung<8, 20, 16, 8, 4, 4>("G16 R8  K20 M4 (T=128)", b, 100);
            b.free_all();
        }
    } else if (set == 7) {      // round 2: four CTAs per SM (16 warps) without the shared-memory halo buffers
        Bufs b; b.alloc(42624, 300, 1.0, 100);
        rung<19, 20, 16, 8, 4, 3>("G16 R19 M3 (12 warps)", b, 100, true);
        rung<19, 20, 16, 8, 4, 4>("G16 R19 M4 plain", b, 100);
        rung<19, 20, 16, 8, 4, 4, 2, 0>("G16 R19 M4 LEAN2 (dy in smem)", b, 100);
        rung<19, 20, 16, 8, 4, 4, 1, 0>("G16 R19 M4 LEAN1 (dy, taps in smem)", b, 100);
        rung<19, 20, 16, 8, 2, 8, 2, 0>("G16 R19 W2 M8 LEAN2", b, 100);
        rung<19, 20, 16, 8, 4, 5, 1, 0>("G16 R19 M5 LEAN1 (20 warps)", b, 100);
        b.free_all();
    } else if (set == 6) {      // round 2: launch bounds of the short-series / 28-tap group variants
        {
            Bufs b; b.alloc(28416, 128, 0.72, 100);
            rung<16, 28, 8, 16, 4, 3>("G8 R16 K28 M3 (T=128)", b, 100, true);
            rung<16, 28, 8, 16, 4, 2>("G8 R16 K28 M2 (T=128)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(28416, 190, 0.72, 100);
            rung<24, 28, 8, 24, 4, 3>("G8 R24 K28 M3 (T=190)", b, 100, true);
            rung<24, 28, 8, 24, 4, 2>("G8 R24 K28 M2 (T=190)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(21312, 240, 0.75, 100);
            rung<15, 28, 16, 8, 4, 3>("G16 R15 K28 M3 (T=240)", b, 100, true);
            rung<15, 28, 16, 8, 4, 2>("G16 R15 K28 M2 (T=240)", b, 100);
            b.free_all();
        }
        {
            Bufs b; b.alloc(21312, 300, 0.72, 100);
            rung<19, 28, 16, 19, 4, 3>("G16 R19 K28 M3 (T=300)", b, 100, true);
            rung<19, 28, 16, 19, 4, 2>("G16 R19 K28 M2 (T=300)", b, 100);
            b.free_all();
        }
    } else if (set == 2) {
        Bufs b; b.alloc(16000, 600, 1.0, 100);
        run32<20, 20, true, 4, 3>("G32 R20 K20 circ W4 M3 (ref)", b, 100, true);
        runc<20, 20, 1, 8>("CTA R20 K20 NW1 M8", b, 100);
        runc<20, 20, 1, 12>("CTA R20 K20 NW1 M12", b, 100);
        runc<20, 20, 1, 16>("CTA R20 K20 NW1 M16", b, 100);
        b.free_all();
    } else {
        Bufs b; b.alloc(8000, 1200, 0.72, 100);
        run32<40, 28, true, 8, 1>("G32 R40 K28 circ W8 M1 (ref)", b, 100, true);
        runc<20, 28, 2, 6>("CTA R20 K28 NW2 M6", b, 100);
        runc<20, 28, 2, 7>("CTA R20 K28 NW2 M7", b, 100);
        runc<20, 28, 2, 8>("CTA R20 K28 NW2 M8", b, 100);
        runc<12, 28, 4, 4>("CTA R12 K28 NW4 M4", b, 100);
        b.free_all();
    }
    return 0;
}
