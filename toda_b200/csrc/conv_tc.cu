// K5/K7 bf16 tensor-core path (tcgen05 + TMEM).  Placeholder until the tcgen05 kernels land:
// reports "unsupported" so callers fail loudly rather than silently taking another path.
#include "common.cuh"

bool conv_tc_supported(int cin, int cout) { (void)cin; (void)cout; return false; }

int conv_tc_fwd(const float *, int, int, const int32_t *, int, int, const float *, int, const float *, float *, cudaStream_t) {
    toda_set_error("conv_tc_fwd: not built");
    return TODA_ERR_UNSUPPORTED;
}
int conv_tc_wgrad(const float *, int, int, const int32_t *, int, int, const float *, int, float *, void *, size_t, cudaStream_t) {
    toda_set_error("conv_tc_wgrad: not built");
    return TODA_ERR_UNSUPPORTED;
}
