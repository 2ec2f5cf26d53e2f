#include "pb_fast_registry.h"

namespace pb {
int fast_deconv_dispatch(const DeconvArgs<float> &, cudaStream_t) { return FAST_NO_MATCH; }
int fast_deconv_dispatch(const DeconvArgs<double> &, cudaStream_t) { return FAST_NO_MATCH; }
int fast_bd_dispatch(const BdArgs<float> &, cudaStream_t) { return FAST_NO_MATCH; }
int fast_bd_dispatch(const BdArgs<double> &, cudaStream_t) { return FAST_NO_MATCH; }
int fast_variant_id(int, int, bool) { return 0; }
}  // namespace pb
