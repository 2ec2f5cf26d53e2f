// Dispatch over the register-tiled instantiations listed in pb_fast_table.inc.
#include "pb_fast_registry.h"

namespace pb {

template <typename real>
struct FastEntry {
    int R, KMAX;
    bool circ;
    bool (*ok)(int T, int K);
    int (*deconv)(const DeconvArgs<real> &, cudaStream_t);
    int (*bd)(const BdArgs<real> &, cudaStream_t);
};

template <typename real>
const FastEntry<real> *fast_table(int *n);

#include "pb_fast_table.inc"

template <typename real>
static const FastEntry<real> *pick(int T, int K) {
    int n = 0;
    const FastEntry<real> *t = fast_table<real>(&n);
    // cheapest matching variant: unrolled work per iteration ~ R * KMAX (the CIRC variants skip
    // the halo / tail selects, hence the small bonus)
    const FastEntry<real> *best = nullptr;
    int best_cost = 0;
    for (int i = 0; i < n; ++i) {
        if (!t[i].ok(T, K)) continue;
        const int cost = t[i].R * t[i].KMAX * (t[i].circ ? 19 : 20);
        if (!best || cost < best_cost) {
            best = &t[i];
            best_cost = cost;
        }
    }
    return best;
}

int fast_deconv_dispatch(const DeconvArgs<float> &a, cudaStream_t s) {
    const FastEntry<float> *e = pick<float>(a.T, a.K);
    return e ? e->deconv(a, s) : FAST_NO_MATCH;
}
int fast_deconv_dispatch(const DeconvArgs<double> &a, cudaStream_t s) {
    const FastEntry<double> *e = pick<double>(a.T, a.K);
    return e ? e->deconv(a, s) : FAST_NO_MATCH;
}
int fast_bd_dispatch(const BdArgs<float> &a, cudaStream_t s) {
    const FastEntry<float> *e = pick<float>(a.T, a.K);
    return e ? e->bd(a, s) : FAST_NO_MATCH;
}
int fast_bd_dispatch(const BdArgs<double> &a, cudaStream_t s) {
    const FastEntry<double> *e = pick<double>(a.T, a.K);
    return e ? e->bd(a, s) : FAST_NO_MATCH;
}
int fast_variant_id(int T, int K, bool is_f64) {
    if (is_f64) {
        const FastEntry<double> *e = pick<double>(T, K);
        return e ? e->R * 1000 + e->KMAX : 0;
    }
    const FastEntry<float> *e = pick<float>(T, K);
    return e ? e->R * 1000 + e->KMAX : 0;
}

}  // namespace pb
