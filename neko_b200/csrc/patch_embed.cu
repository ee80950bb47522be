// Image patch embedding (placeholder until the fused kernels land in this file).
#include "common.cuh"

extern "C" {

int neko_patch_resblock_fwd(const void*, int, int, int, int, int, int, int, const float*, const float*, const float*, const float*,
                            const float*, const float*, uint16_t*, float*, void*) {
  neko::set_error("patch_resblock_fwd: not implemented yet");
  return NEKO_EINVAL;
}
int neko_patch_resblock_bwd(const void*, int, int, int, int, int, int, int, const float*, const float*, const float*, const float*,
                            const float*, const float*, const uint16_t*, float*, float*, float*, float*, float*, float*, void*) {
  neko::set_error("patch_resblock_bwd: not implemented yet");
  return NEKO_EINVAL;
}
int neko_patch_pos_add(float*, int, int, int, int, const int32_t*, const int32_t*, const float*, const float*, void*) {
  neko::set_error("patch_pos_add: not implemented yet");
  return NEKO_EINVAL;
}
int neko_patch_pos_bwd(const float*, int, int, int, int, const int32_t*, const int32_t*, float*, float*, void*) {
  neko::set_error("patch_pos_bwd: not implemented yet");
  return NEKO_EINVAL;
}

}  // extern "C"
