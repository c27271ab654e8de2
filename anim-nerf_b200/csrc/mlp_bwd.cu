// placeholder: replaced by the tcgen05 backward (dgrad chain + wgrad)
#include "common.cuh"
#include "mlp_layout.cuh"
extern "C" int64_t an_mlp_bwd_scratch_bytes(int64_t n_max) { return n_max > 0 ? 16 : 0; }
extern "C" int an_mlp_bwd(const void*, const void*, const float*, const int32_t*, const int32_t*, int64_t,
                          const float*, const float*, float*, float*, void*, void*) { return AN_ERR_UNSUPPORTED; }
