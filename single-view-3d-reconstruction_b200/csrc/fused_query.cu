// Fused gather -> tcgen05 decoder kernel and the dense-grid evaluator (placeholder translation unit:
// the entry points exist so that the library matches include/svr_b200.h; they report an error until
// the fused kernel lands).
#include "common.cuh"
#include "sampling.cuh"

extern "C" {

int svr_query_fwd_fused(const float *, int, int, const float *, const uint16_t *const *, const svr_pyramid *,
                        const svr_decoder_weights *, float *, uint16_t *, uint16_t *, int, void *) {
    svr::set_error("svr_query_fwd_fused: not available in this build");
    return -2;
}

int svr_dense_eval(int, int, const float *, const uint16_t *const *, const svr_pyramid *, const svr_decoder_weights *, int,
                   int, int, int, int, float *, void *) {
    svr::set_error("svr_dense_eval: not available in this build");
    return -2;
}
}
