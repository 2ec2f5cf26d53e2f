// Dispatch table of the register-tiled warp solver instantiations (pb_fast.cuh).  Each
// instantiation lives in its own translation unit (pb_fast_inst_*.cu) so that they compile in
// parallel; pb_fast_registry.cu picks one from (T, K, dtype).
#pragma once
#include <cuda_runtime.h>

#include "pb_generic.cuh"

namespace pb {

constexpr int FAST_NO_MATCH = -1000;

int fast_deconv_dispatch(const DeconvArgs<float> &a, cudaStream_t stream);
int fast_deconv_dispatch(const DeconvArgs<double> &a, cudaStream_t stream);
int fast_bd_dispatch(const BdArgs<float> &a, cudaStream_t stream);
int fast_bd_dispatch(const BdArgs<double> &a, cudaStream_t stream);
// 0 = generic kernel, else R * 1000 + KMAX
int fast_variant_id(int T, int K, bool is_f64);
// voxels per full wave of the persistent bd grid (0 when unknown): batch sizes that are multiples
// of it waste no tail
int fast_bd_wave_voxels(int T, int K, bool is_f64, int nb_iter);

}  // namespace pb
