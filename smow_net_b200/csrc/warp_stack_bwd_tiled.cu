// A1 backward, tiled inverse-gather variant (placeholder until the kernel lands).
#include "warp_stack_tiled.cuh"

namespace smow {

template <typename T>
int warp_bwd_tiled(const T*, const T*, const T*, int64_t, int64_t, const float*, const float*, const float*, T*,
                   T*, float*, int, int, int, int, cudaStream_t) {
  return fail(SMOW_EINVAL, "warp_bwd_variant 1 is not built in this revision");
}
template int warp_bwd_tiled<float>(const float*, const float*, const float*, int64_t, int64_t, const float*,
                                   const float*, const float*, float*, float*, float*, int, int, int, int,
                                   cudaStream_t);
template int warp_bwd_tiled<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*,
                                           int64_t, int64_t, const float*, const float*, const float*,
                                           __nv_bfloat16*, __nv_bfloat16*, float*, int, int, int, int,
                                           cudaStream_t);
}  // namespace smow
