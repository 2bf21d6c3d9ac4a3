// coarse_tc.cu -- placeholder until the tcgen05 kernel lands (next commit).
#include "common.cuh"
namespace rvcr {
bool coarse_tc_available() { return false; }
int coarse_tc_invalidate(rvcr_index*) { return RVCR_OK; }
int coarse_tc(rvcr_index*, const float*, int64_t, int32_t*, float*, cudaStream_t) {
    set_error("tcgen05 coarse path not built");
    return RVCR_ERR_UNSUPPORTED;
}
}  // namespace rvcr
