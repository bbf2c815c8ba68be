// FA loss, POSITION semantics (placeholder until the tcgen05 tile engine lands).
#include "common.cuh"

namespace dsrl {
size_t fa_pos_saved_bytes(int, int, int, int, int, int) { return 0; }
size_t fa_pos_workspace_bytes(int, int, int, int, int, int) { return 0; }
int fa_pos_forward(int, const float *, const float *, int, int, int, int, int, int, int, int, float *, void *, size_t, void *, size_t, cudaStream_t) {
    DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(position): not built yet");
}
int fa_pos_backward(int, const float *, const float *, const void *, size_t, const float *, float *, float *, int, int, int, int, int, int, int, void *, size_t, cudaStream_t) {
    DSRL_FAIL(DSRL_ERR_UNSUPPORTED, "FA(position): not built yet");
}
}  // namespace dsrl
