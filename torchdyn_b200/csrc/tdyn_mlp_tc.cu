// K2 placeholder until the tcgen05 kernel lands: reports "unsupported" so callers use the generic path.
#include "tdyn_common.cuh"
extern "C" {
int64_t tdyn_mlp_tc_prepared_bytes(const tdyn_mlp_t*) { return 0; }
int tdyn_mlp_tc_supported(const tdyn_mlp_t*) { return 0; }
int tdyn_mlp_tc_prepare(const tdyn_mlp_t*, void*, void*) { tdyn::set_error("tdyn_mlp_tc: not built"); return 3; }
int tdyn_mlp_tc_stage(const tdyn_mlp_t*, const void*, float*, float*, const float*, const float* const*, const float*, int,
                      int, float, float*, const float*, int, int64_t, void*) { tdyn::set_error("tdyn_mlp_tc: not built"); return 3; }
}
