// A1, channels-last (NDHWC) variants (placeholder until the kernels land).
#include "warp_stack_tiled.cuh"

namespace smow {

template <typename T>
int warp_fwd_ndhwc(const T*, const T*, int64_t, const float*, const float*, const float*, T*, int, int, int, int,
                   cudaStream_t) {
  return fail(SMOW_EDTYPE, "NDHWC warp kernels are not built in this revision");
}
template <typename T>
int warp_bwd_ndhwc(const T*, const T*, const T*, int64_t, const float*, const float*, const float*, T*, T*,
                   float*, int, int, int, int, cudaStream_t) {
  return fail(SMOW_EDTYPE, "NDHWC warp kernels are not built in this revision");
}
#define INST(T)                                                                                              \
  template int warp_fwd_ndhwc<T>(const T*, const T*, int64_t, const float*, const float*, const float*, T*, \
                                 int, int, int, int, cudaStream_t);                                         \
  template int warp_bwd_ndhwc<T>(const T*, const T*, const T*, int64_t, const float*, const float*,         \
                                 const float*, T*, T*, float*, int, int, int, int, cudaStream_t);
INST(float)
INST(__nv_bfloat16)
}  // namespace smow
