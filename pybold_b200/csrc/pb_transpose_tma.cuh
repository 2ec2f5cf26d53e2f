// Row N4: [rows, cols] -> [cols, rows] transposition (the reference pipeline's time-major [T, V] voxel
// matrices, examples/icassp_2019/validation.py:90-93,103, to the solvers' [V, T] layout and back) as a
// pure tile-movement kernel on the sm_100a copy hardware:
//   TMA loads  (cp.async.bulk.tensor.2d ... mbarrier::complete_tx) bring E x E boxes of the source into
//   shared memory, E = 128 bytes / sizeof(real), with the 128-byte swizzle, STAGES tiles in flight per CTA;
//   the warps move every box to its transposed place: conflict-free scalar reads along a source row, one
//   16-byte store per lane into the (swizzled) destination box;
//   TMA stores (cp.async.bulk.tensor.2d.global.shared::cta.bulk_group) write the boxes out.
// No thread ever forms a global address; out-of-range parts of edge boxes are zero-filled on load and
// clipped on store by the tensor maps.  Persistent grid, tiles strided over the CTAs.
// Needs 16-byte aligned pointers and row pitches (cols and rows multiples of 16 / sizeof(real));
// other shapes stay on pb::transpose_kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace pb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "PB_MBAR_WAIT:\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "  @p bra PB_MBAR_DONE;\n"
        "  bra PB_MBAR_WAIT;\n"
        "PB_MBAR_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(src)
                 : "memory");
}

template <typename real>
struct TransposeTma {
    static constexpr int ES = (int)sizeof(real);
    static constexpr int E = 128 / ES;             // box side in elements (one 128-byte swizzle row)
    static constexpr int BOX_BYTES = E * 128;      // 4 KB (float) / 2 KB (double), 1024-byte aligned
    static constexpr int BX = ES == 4 ? 2 : 4;     // boxes per tile side: 64 x 64 elements either way
    static constexpr int TILE = BX * E;
    static constexpr int TILE_BYTES = BX * BX * BOX_BYTES;
    static constexpr int STAGES = ES == 4 ? 4 : 3;
    static constexpr int THREADS = 256;
    static constexpr int PER16 = 16 / ES;          // elements per 16-byte chunk
    static constexpr size_t SMEM = (size_t)(STAGES + 2) * TILE_BYTES + 1024 /* alignment slack */ + 64;
};

// byte offset of element (r, c) inside a 128B-swizzled box whose rows are 128 bytes
template <int ES>
__device__ __forceinline__ int swz_off(int r, int c) {
    const int byte = c * ES;
    return r * 128 + (((byte >> 4) ^ (r & 7)) << 4) + (byte & 15);
}

template <typename real>
__global__ void __launch_bounds__(256)
transpose_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                     int tiles_r, int tiles_c) {
    using C = TransposeTma<real>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle, as an offset into the __shared__ array so that the
    // compiler keeps the shared address space (LDS / STS, not generic loads)
    unsigned char *base = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char *in_s = base;                                    // STAGES tiles
    unsigned char *out_s = base + (size_t)C::STAGES * C::TILE_BYTES;   // 2 tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(out_s + 2 * C::TILE_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_tiles = (int64_t)tiles_r * tiles_c;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t my_tiles = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;

    // Tile order: the index along the dimension with FEWER tiles runs fastest, so that the tiles in flight at
    // any time (CTAs x stages) cover whole rows of the narrow side: e.g. [T, V] -> [V, T] with T = 1200 writes
    // each voxel row (4.8 KB) from 19 consecutive tiles instead of 256-byte pieces 3594 tiles apart.
    const bool r_fast = tiles_r < tiles_c;
    auto tile_rc = [&](int64_t t, int &tr, int &tc) {
        if (r_fast) {
            tc = (int)(t / tiles_r);
            tr = (int)(t % tiles_r);
        } else {
            tr = (int)(t / tiles_c);
            tc = (int)(t % tiles_c);
        }
    };
    auto issue_load = [&](int64_t it) {            // tile number `it` of this CTA into stage it % STAGES
        const int64_t t = first + it * stride;
        int tr, tc;
        tile_rc(t, tr, tc);
        const int s = (int)(it % C::STAGES);
        const uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, C::TILE_BYTES);
#pragma unroll
        for (int bi = 0; bi < C::BX; ++bi)
#pragma unroll
            for (int bj = 0; bj < C::BX; ++bj)
                tma_load_2d(smem_u32(in_s + (size_t)s * C::TILE_BYTES + (bi * C::BX + bj) * C::BOX_BYTES), &in_map,
                            tc * C::TILE + bj * C::E, tr * C::TILE + bi * C::E, bar);
    };

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int64_t it = 0; it < C::STAGES && it < my_tiles; ++it) issue_load(it);

    for (int64_t it = 0; it < my_tiles; ++it) {
        const int s = (int)(it % C::STAGES);
        const uint32_t parity = (uint32_t)((it / C::STAGES) & 1);
        unsigned char *ob = out_s + (size_t)(it & 1) * C::TILE_BYTES;
        // the bulk store issued two tiles ago from this output buffer must have read it
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        mbar_wait(smem_u32(&bars[s]), parity);
        const unsigned char *ib = in_s + (size_t)s * C::TILE_BYTES;
        // unit = (box bi, bj; block of PER16 source rows); a lane takes column c of the box
        constexpr int LANES_PER_UNIT = C::E;                       // 32 (float) / 16 (double)
        constexpr int UNITS_PER_WARP_OP = 32 / LANES_PER_UNIT;     // 1 / 2
        constexpr int RBLOCKS = C::E / C::PER16;                   // 8 row blocks per box
        constexpr int UNITS = C::BX * C::BX * RBLOCKS;
        const int c = lane % LANES_PER_UNIT, sub = lane / LANES_PER_UNIT;
#pragma unroll
        for (int u0 = 0; u0 < UNITS; u0 += 8 * UNITS_PER_WARP_OP) {
            const int u = u0 + warp * UNITS_PER_WARP_OP + sub;
            const int box = u / RBLOCKS, rb = u % RBLOCKS;
            const int bi = box / C::BX, bj = box % C::BX;
            const unsigned char *src = ib + (bi * C::BX + bj) * C::BOX_BYTES;
            unsigned char *dst = ob + (bj * C::BX + bi) * C::BOX_BYTES;   // transposed box position
            const int r0 = rb * C::PER16;
            struct alignas(16) Chunk { real v[C::PER16]; } ch;
#pragma unroll
            for (int i = 0; i < C::PER16; ++i)
                ch.v[i] = *reinterpret_cast<const real *>(src + swz_off<C::ES>(r0 + i, c));
            // destination box: row = source column c, 16-byte chunk = source rows r0 .. r0 + PER16 - 1
            *reinterpret_cast<Chunk *>(dst + swz_off<C::ES>(c, r0)) = ch;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> TMA store
        __syncthreads();
        if (tid == 0) {
            const int64_t t = first + it * stride;
            int tr, tc;
            tile_rc(t, tr, tc);
#pragma unroll
            for (int bj = 0; bj < C::BX; ++bj)
#pragma unroll
                for (int bi = 0; bi < C::BX; ++bi)
                    tma_store_2d(&out_map, tr * C::TILE + bi * C::E, tc * C::TILE + bj * C::E,
                                 smem_u32(ob + (bj * C::BX + bi) * C::BOX_BYTES));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (it + C::STAGES < my_tiles) issue_load(it + C::STAGES);   // stage s is free again
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// map of a row-major [outer, inner] matrix with E x E boxes and the 128-byte swizzle
template <typename real>
bool make_matrix_map(CUtensorMap *map, const real *ptr, int64_t outer, int64_t inner) {
    using C = TransposeTma<real>;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)inner * sizeof(real)};
    const cuuint32_t box[2] = {(cuuint32_t)C::E, (cuuint32_t)C::E};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = sizeof(real) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    return enc(map, dt, 2, const_cast<real *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns cudaSuccess (0), a CUDA error, or -1000 when the shape / alignment is not TMA-friendly
template <typename real>
int transpose_tma_launch(const real *in, real *out, int64_t rows, int64_t cols, int sm_count, cudaStream_t stream) {
    using C = TransposeTma<real>;
    constexpr int NO = -1000;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return NO;
    if ((cols * (int64_t)sizeof(real)) % 16 || (rows * (int64_t)sizeof(real)) % 16) return NO;
    if (rows >= (1LL << 31) || cols >= (1LL << 31)) return NO;
    CUtensorMap in_map, out_map;
    if (!make_matrix_map<real>(&in_map, in, rows, cols) || !make_matrix_map<real>(&out_map, out, cols, rows)) return NO;
    const int64_t tr = (rows + C::TILE - 1) / C::TILE, tc = (cols + C::TILE - 1) / C::TILE;
    if (tr >= (1LL << 31) || tc >= (1LL << 31)) return NO;
    auto kern = transpose_tma_kernel<real>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return (int)e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    if (occ < 1) return NO;
    const int64_t cap = (int64_t)sm_count * occ, n = tr * tc;
    kern<<<(int)(n < cap ? n : cap), C::THREADS, C::SMEM, stream>>>(in_map, out_map, (int)tr, (int)tc);
    e = cudaGetLastError();
    return (int)e;
}

}  // namespace pb
