// Shared device/host helpers for libsmow_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/smow_b200.h"

namespace smow {

// ---- host side: error slot, launch counter, tuning knobs (abi.cu) -------------
int  fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int  option(int id);
enum OptionId { OPT_WARP_FWD_VARIANT = 0, OPT_WARP_BWD_VARIANT, OPT_TLERP_VARIANT,
                OPT_BWD_ROWS, OPT_BWD_HALO, OPT_FWD_ROWS, OPT_FWD_HALO, OPT_CVEC_PF, OPT_BWD_CHUNK_MB, OPT_NDHWC_BWD_ROWS, OPT_NDHWC_BWD_PF, OPT_TC_DEBUG, OPT_TOK_VARIANT, OPT_BN_BWD_ROWS, OPT_COUNT };

struct DeviceInfo { int sms; int smem_optin; };
DeviceInfo device_info();   // cached per device (abi.cu)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- element access: features are fp32 or bf16, arithmetic is always fp32 -----
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <typename T> __device__ __forceinline__ float cvtf(T v);
template <> __device__ __forceinline__ float cvtf<float>(float v) { return v; }
template <> __device__ __forceinline__ float cvtf<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T fromf(float v);
template <> __device__ __forceinline__ float fromf<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 fromf<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- the reference's fp32 coordinate chain, bit for bit -----------------------
// models/SMOW_Net.py:627-631:  g = base + flow / size ; g = clamp(g, -1, 1)
// ATen GridSampler.cuh:21-31 :  i = ((g + 1) / 2) * (size - 1)       (align_corners)
// ATen GridSampler.cuh:53-80 :  i = min(size-1, max(i, 0)); d(clip)/di = 0 at i<=0 or i>=size-1
// Every step is an explicitly rounded fp32 op so the compiler cannot contract it.
struct Axis {
  float i;      // clipped source coordinate in pixels, in [0, size-1]
  float i0f;    // floor(i)
  int   i0;     // (int)floor(i)
  float gmult;  // d(i)/d(flow) including clamp mask, unnormalise and clip masks (0 where blocked)
};

__device__ __forceinline__ Axis axis_coord(float base, float flow, int size) {
  const float sz = (float)size;
  float g = __fadd_rn(base, __fdiv_rn(flow, sz));
  const bool pass_clamp = (g >= -1.f) && (g <= 1.f);     // ClampBackward mask, inclusive
  g = fminf(fmaxf(g, -1.f), 1.f);
  float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  const float hi = (float)(size - 1);
  const bool pass_clip = (i > 0.f) && (i < hi);           // clip_coordinates_set_grad
  i = fminf(hi, fmaxf(i, 0.f));
  Axis a;
  a.i = i;
  a.i0f = floorf(i);
  a.i0 = (int)a.i0f;
  // grad chain: clip(0/1) * (size-1)/2 [unnormalise] * clamp mask * 1/size [DivBackward]
  // applied as  ((gacc * ((size-1)/2)) / size)  in kernels; here only the 0/1 gate.
  a.gmult = (pass_clamp && pass_clip) ? 1.f : 0.f;
  return a;
}

// Bilinear footprint of one output pixel: 4 weights in ATen's order nw, ne, sw, se
// (GridSampler.cu: nw = (ix_se-ix)*(iy_se-iy) ...), tap validity, and the partial
// derivatives' factors.
struct Footprint {
  int   x0, y0;        // nw corner
  float wx0, wx1;      // (x0+1 - ix), (ix - x0)
  float wy0, wy1;      // (y0+1 - iy), (iy - y0)
  bool  x1ok, y1ok;    // x0+1 <= W-1, y0+1 <= H-1   (x0,y0 always in bounds after the clip)
  float gx_gate, gy_gate;
};

__device__ __forceinline__ Footprint footprint(float xs_w, float ys_h, float fx, float fy,
                                               int W, int H) {
  Axis ax = axis_coord(xs_w, fx, W);
  Axis ay = axis_coord(ys_h, fy, H);
  Footprint f;
  f.x0 = ax.i0; f.y0 = ay.i0;
  f.wx0 = __fsub_rn(__fadd_rn(ax.i0f, 1.f), ax.i);
  f.wx1 = __fsub_rn(ax.i, ax.i0f);
  f.wy0 = __fsub_rn(__fadd_rn(ay.i0f, 1.f), ay.i);
  f.wy1 = __fsub_rn(ay.i, ay.i0f);
  f.x1ok = (ax.i0 + 1) <= (W - 1);
  f.y1ok = (ay.i0 + 1) <= (H - 1);
  f.gx_gate = ax.gmult; f.gy_gate = ay.gmult;
  return f;
}

// Same chain with the division by a power-of-two size done as an exact multiplication by 2^-k
// (bit-identical result, ~20 instructions shorter than the IEEE division sequence).
__device__ __forceinline__ Axis axis_coord_pow2(float base, float flow, int size, float inv) {
  float g = __fadd_rn(base, __fmul_rn(flow, inv));
  const bool pass_clamp = (g >= -1.f) && (g <= 1.f);
  g = fminf(fmaxf(g, -1.f), 1.f);
  const float hi = (float)(size - 1);
  float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), hi);
  const bool pass_clip = (i > 0.f) && (i < hi);
  i = fminf(hi, fmaxf(i, 0.f));
  Axis a;
  a.i = i; a.i0f = floorf(i); a.i0 = (int)a.i0f;
  a.gmult = (pass_clamp && pass_clip) ? 1.f : 0.f;
  return a;
}

// footprint() with the power-of-two shortcut taken per axis when the size allows it (uniform branch)
__device__ __forceinline__ Footprint footprint_auto(float xs_w, float ys_h, float fx, float fy, int W, int H) {
  const bool wp = (W & (W - 1)) == 0, hp = (H & (H - 1)) == 0;
  const Axis ax = wp ? axis_coord_pow2(xs_w, fx, W, 1.f / (float)W) : axis_coord(xs_w, fx, W);
  const Axis ay = hp ? axis_coord_pow2(ys_h, fy, H, 1.f / (float)H) : axis_coord(ys_h, fy, H);
  Footprint f;
  f.x0 = ax.i0; f.y0 = ay.i0;
  f.wx0 = __fsub_rn(__fadd_rn(ax.i0f, 1.f), ax.i);
  f.wx1 = __fsub_rn(ax.i, ax.i0f);
  f.wy0 = __fsub_rn(__fadd_rn(ay.i0f, 1.f), ay.i);
  f.wy1 = __fsub_rn(ay.i, ay.i0f);
  f.x1ok = (ax.i0 + 1) <= (W - 1);
  f.y1ok = (ay.i0 + 1) <= (H - 1);
  f.gx_gate = ax.gmult; f.gy_gate = ay.gmult;
  return f;
}

// temporal lerp weights of upsample_trilinear3d(2 -> 4, align_corners=True):
// rdepth = (2-1)/(4-1) as fp32; t1lambda = rdepth * t2 - floor(.)
#define SMOW_LAMBDA1 0.3333333432674408f
#define SMOW_LAMBDA2 0.6666666865348816f

}  // namespace smow
