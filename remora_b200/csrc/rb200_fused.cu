// placeholder, replaced below
#include "rb200_internal.cuh"
namespace rb200 {
struct FusedWeights {};
bool fused_supported(const rb200_model_desc &) { return false; }
int fused_create(rb200_model *, const float *) { return RB200_OK; }
void fused_destroy(rb200_model *) {}
bool fused_shape_ok(const rb200_model *, int, int, int) { return false; }
size_t fused_workspace_bytes(const rb200_model *, int, int) { return 0; }
int fused_forward_compact(rb200_model *, Workspace &, const float *, const int8_t *, int,
                          const int16_t *, int, const int16_t *, int, int, float *, cudaStream_t) {
    return RB200_ERR_UNSUPPORTED;
}
}  // namespace rb200
