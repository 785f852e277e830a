// A1: fused optical-flow warp + temporal stack, forward and backward (sm_100a).
//
// Reference graph replaced (models/SMOW_Net.py:612-638):
//   fwd: out = cat([x[:,:,0], grid_sample(x[:,:,0], g0), grid_sample(x[:,:,1], g1), x[:,:,1]], dim=2)
//   bwd: gx[:,:,t] = gout[:,:,pass(t)] + scatter_t(gout[:,:,1+t]);  gflow = d out / d flow
//
// Variant 0 ("direct"): one thread per output pixel, channels looped, taps fetched
// through L1 (what ATen does, minus its 30 helper launches); backward scatters with
// L2 atomics.  It is the simple cross-check for the tiled variants in
// warp_stack_tiled.cuh (bulk-copy staged tiles, inverse-gather backward).
#include <type_traits>
#include "common.cuh"
#include "warp_stack_tiled.cuh"

namespace smow {

// Bi-temporal frame addressing, NCDHW: element (b,c,t,h,w) = f[t][b*sB + c*sC + h*W + w].
// Stacked (B,C,2,H,W): f[1] = f[0] + HW, sC = 2HW, sB = 2*C*HW.  Pair of (B,C,H,W): sC = HW, sB = C*HW.
template <typename T> struct Frames {
  T* f0; T* f1; int64_t sB, sC;
  __device__ __forceinline__ T* frame(int t) const { return t ? f1 : f0; }
};

// ------------------------------------------------------------------------------
// variant 0 forward
// ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
warp_stack_fwd_direct(Frames<const T> x, const float* __restrict__ flow,
                      const float* __restrict__ xs, const float* __restrict__ ys,
                      T* __restrict__ out, int C, int H, int W) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  const int h = p / W, w = p - h * W;
  const float fx = __ldg(flow + ((int64_t)(b * 2 + 0) * 2 + t) * HW + p);
  const float fy = __ldg(flow + ((int64_t)(b * 2 + 1) * 2 + t) * HW + p);
  const Footprint fp = footprint(__ldg(xs + w), __ldg(ys + h), fx, fy, W, H);
  const float nw = __fmul_rn(fp.wx0, fp.wy0), ne = __fmul_rn(fp.wx1, fp.wy0);
  const float sw = __fmul_rn(fp.wx0, fp.wy1), se = __fmul_rn(fp.wx1, fp.wy1);
  const int o_nw = fp.y0 * W + fp.x0;
  const int o_ne = o_nw + (fp.x1ok ? 1 : 0);
  const int o_sw = o_nw + (fp.y1ok ? W : 0);
  const int o_se = o_sw + (fp.x1ok ? 1 : 0);
  const float m_ne = fp.x1ok ? 1.f : 0.f, m_sw = fp.y1ok ? 1.f : 0.f;
  const float m_se = (fp.x1ok && fp.y1ok) ? 1.f : 0.f;
  const T* src = x.frame(t) + b * x.sB;
  T* dst_warp = out + ((int64_t)b * C * 4 + (1 + t)) * HW + p;
  T* dst_pass = out + ((int64_t)b * C * 4 + (t ? 3 : 0)) * HW + p;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const T* pl = src + c * x.sC;
    // ATen order: nw, ne, sw, se; an out-of-bounds tap is skipped, not multiplied by 0
    float acc = __fmul_rn(ldf(pl + o_nw), nw);
    float v_ne = ldf(pl + o_ne), v_sw = ldf(pl + o_sw), v_se = ldf(pl + o_se);
    if (m_ne != 0.f) acc = fmaf(v_ne, ne, acc);
    if (m_sw != 0.f) acc = fmaf(v_sw, sw, acc);
    if (m_se != 0.f) acc = fmaf(v_se, se, acc);
    dst_warp[(int64_t)c * 4 * HW] = fromf<T>(acc);
    dst_pass[(int64_t)c * 4 * HW] = pl[p];
  }
}

// ------------------------------------------------------------------------------
// variant 0 backward: (1) gx = pass-through slots, (2) scatter + flow gradient
// ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
warp_stack_bwd_pass(const T* __restrict__ gout, Frames<T> gx, int C, int HW) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int bc = blockIdx.y;  // b*C + c
  const int b = bc / C, c = bc - b * C;
#pragma unroll
  for (int t = 0; t < 2; ++t)
    gx.frame(t)[b * gx.sB + c * gx.sC + p] = gout[((int64_t)bc * 4 + (t ? 3 : 0)) * HW + p];
}

template <typename T> __device__ __forceinline__ void atomic_addf(T* p, float v);
template <> __device__ __forceinline__ void atomic_addf<float>(float* p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void atomic_addf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

template <typename T>
__global__ void __launch_bounds__(256)
warp_stack_bwd_scatter(const T* __restrict__ gout, Frames<const T> x, const float* __restrict__ flow,
                       const float* __restrict__ xs, const float* __restrict__ ys,
                       Frames<T> gx, float* __restrict__ gflow, int C, int H, int W) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  const int h = p / W, w = p - h * W;
  const int64_t fo = ((int64_t)(b * 2 + 0) * 2 + t) * HW + p;
  const float fx = __ldg(flow + fo), fy = __ldg(flow + fo + 2 * (int64_t)HW);
  const Footprint fp = footprint(__ldg(xs + w), __ldg(ys + h), fx, fy, W, H);
  const float nw = __fmul_rn(fp.wx0, fp.wy0), ne = __fmul_rn(fp.wx1, fp.wy0);
  const float sw = __fmul_rn(fp.wx0, fp.wy1), se = __fmul_rn(fp.wx1, fp.wy1);
  const int o_nw = fp.y0 * W + fp.x0;
  const T* src = x.frame(t) + b * x.sB;
  T* dst = gx.frame(t) + b * gx.sB;
  const T* g = gout + ((int64_t)b * C * 4 + (1 + t)) * HW + p;
  float gix = 0.f, giy = 0.f;
  for (int c = 0; c < C; ++c) {
    const float go = cvtf<T>(g[(int64_t)c * 4 * HW]);
    const T* pl = src + c * x.sC;
    T* gp = dst + c * gx.sC;
    {
      atomic_addf(gp + o_nw, __fmul_rn(nw, go));
      const float v = ldf(pl + o_nw);
      gix = fmaf(-__fmul_rn(v, fp.wy0), go, gix);
      giy = fmaf(-__fmul_rn(v, fp.wx0), go, giy);
    }
    if (fp.x1ok) {
      atomic_addf(gp + o_nw + 1, __fmul_rn(ne, go));
      const float v = ldf(pl + o_nw + 1);
      gix = fmaf(__fmul_rn(v, fp.wy0), go, gix);
      giy = fmaf(-__fmul_rn(v, fp.wx1), go, giy);
    }
    if (fp.y1ok) {
      atomic_addf(gp + o_nw + W, __fmul_rn(sw, go));
      const float v = ldf(pl + o_nw + W);
      gix = fmaf(-__fmul_rn(v, fp.wy1), go, gix);
      giy = fmaf(__fmul_rn(v, fp.wx0), go, giy);
    }
    if (fp.x1ok && fp.y1ok) {
      atomic_addf(gp + o_nw + W + 1, __fmul_rn(se, go));
      const float v = ldf(pl + o_nw + W + 1);
      gix = fmaf(__fmul_rn(v, fp.wy1), go, gix);
      giy = fmaf(__fmul_rn(v, fp.wx1), go, giy);
    }
  }
  // GridSampler2DBackward * ClampBackward * DivBackward
  const float mx = __fmul_rn(fp.gx_gate, __fmul_rn((float)(W - 1), 0.5f));
  const float my = __fmul_rn(fp.gy_gate, __fmul_rn((float)(H - 1), 0.5f));
  gflow[fo] = __fdiv_rn(__fmul_rn(mx, gix), (float)W);
  gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, giy), (float)H);
}

// ------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------
static int check_common(const void* a, const void* b, const void* c, const void* d,
                        int B, int C, int H, int W, int dtype, int layout) {
  if (!a || !b || !c || !d) return fail(SMOW_EINVAL, "null pointer argument");
  if (B <= 0 || C <= 0 || H <= 1 || W <= 1) return fail(SMOW_EINVAL, "bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  if (dtype != SMOW_F32 && dtype != SMOW_BF16) return fail(SMOW_EDTYPE, "unsupported dtype %d", dtype);
  if (layout != SMOW_NCDHW && layout != SMOW_NDHWC) return fail(SMOW_EDTYPE, "unsupported layout %d", layout);
  if ((int64_t)H * W >= (1 << 24)) return fail(SMOW_ERANGE, "H*W too large");
  if (2 * B > 65535) return fail(SMOW_ERANGE, "B too large for one launch");
  return 0;
}

template <typename T>
static int fwd_impl(const T* x1, const T* x2, int64_t sB, int64_t sC, const float* flow,
                    const float* xs, const float* ys, T* out, int B, int C, int H, int W,
                    int layout, cudaStream_t st) {
  if (layout == SMOW_NDHWC) return warp_fwd_ndhwc<T>(x1, x2, sB, flow, xs, ys, out, B, C, H, W, st);
  int variant = option(OPT_WARP_FWD_VARIANT);
  if (variant < -1 || variant > 2) return fail(SMOW_EINVAL, "unknown warp_fwd_variant %d", variant);
  // auto (-1), from the r1 sweep (profiles/r1_sweep.md): the bulk-copy staged planes win for few
  // channels, the channel-vectorised tiles from 64 channels up
  if (variant == -1) variant = (std::is_same<T, float>::value && C >= 64) ? 2 : 1;
  // the tile kernels need whole 16 B rows; odd shapes take the direct kernel (same results)
  const bool tile_ok = tiled_shape_ok<T>(x1, x2, out, sB, sC, C, H, W);
  if constexpr (std::is_same<T, float>::value) {
    if (variant == 2 && tile_ok) {
      const int rc = warp_fwd_cvec(x1, x2, sB, sC, flow, xs, ys, out, B, C, H, W, st);
      if (rc != SMOW_ERANGE) return rc;
    }
  }
  if (variant >= 1 && tile_ok) return warp_fwd_tiled<T>(x1, x2, sB, sC, flow, xs, ys, out, B, C, H, W, st);
  Frames<const T> x{x1, x2, sB, sC};
  dim3 grid((H * W + 255) / 256, 2 * B);
  warp_stack_fwd_direct<T><<<grid, 256, 0, st>>>(x, flow, xs, ys, out, C, H, W);
  count_launch();
  return check_launch("warp_stack_fwd_direct");
}

template <typename T>
static int bwd_impl(const T* gout, const T* x1, const T* x2, int64_t sB, int64_t sC,
                    const float* flow, const float* xs, const float* ys,
                    T* gx1, T* gx2, float* gflow, int B, int C, int H, int W,
                    int layout, void* ws, int64_t ws_bytes, cudaStream_t st) {
  if (layout == SMOW_NDHWC)
    return warp_bwd_ndhwc<T>(gout, x1, x2, sB, flow, xs, ys, gx1, gx2, gflow, B, C, H, W, ws, ws_bytes, st);
  int variant = option(OPT_WARP_BWD_VARIANT);
  if (variant < -1 || variant > 3) return fail(SMOW_EINVAL, "unknown warp_bwd_variant %d", variant);
  if (variant == -1 || variant == 3) variant = 2;   // NCDHW: the tile gathers are deterministic already; falls
                                                     // through to 1 for bf16 / uncovered shapes, then to 0
  const bool tile_ok = tiled_shape_ok<T>(x1, x2, gout, sB, sC, C, H, W) && aligned16(gx1) && aligned16(gx2);
  if constexpr (std::is_same<T, float>::value) {
    if (variant == 2 && tile_ok) {
      const int rc = warp_bwd_cvec(gout, x1, x2, sB, sC, flow, xs, ys, gx1, gx2, gflow, B, C, H, W, st);
      if (rc != SMOW_ERANGE) return rc;
    }
  }
  if (variant >= 1 && tile_ok) {
    const int rc = warp_bwd_tiled<T>(gout, x1, x2, sB, sC, flow, xs, ys, gx1, gx2, gflow, B, C, H, W, st);
    if (rc != SMOW_ERANGE) return rc;   // rows too wide for one tile: the scatter kernel below handles them
  }
  if ((int64_t)B * C > 65535) return fail(SMOW_ERANGE, "B*C too large for variant 0");
  Frames<const T> x{x1, x2, sB, sC};
  Frames<T> gx{gx1, gx2, sB, sC};
  const int HW = H * W;
  warp_stack_bwd_pass<T><<<dim3((HW + 255) / 256, B * C), 256, 0, st>>>(gout, gx, C, HW);
  warp_stack_bwd_scatter<T><<<dim3((HW + 255) / 256, 2 * B), 256, 0, st>>>(gout, x, flow, xs, ys, gx,
                                                                          gflow, C, H, W);
  count_launch(2);
  return check_launch("warp_stack_bwd (variant 0)");
}

}  // namespace smow

using namespace smow;

extern "C" {

int64_t smow_warp_bwd_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return warp_bwd_ndhwc_workspace_bytes(B, H, W);   // header + coordinates, anchors and gather lists of every pixel-frame
}

int smow_warp_pair_fwd(const void* x_t1, const void* x_t2, const float* flow, const float* xs,
                       const float* ys, void* out, int B, int C, int H, int W, int dtype, int layout,
                       void* stream) {
  if (!x_t2 || !xs || !ys) return fail(SMOW_EINVAL, "null pointer argument");
  if (int e = check_common(x_t1, flow, out, xs, B, C, H, W, dtype, layout)) return e;
  const int64_t HW = (int64_t)H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMOW_F32)
    return fwd_impl<float>((const float*)x_t1, (const float*)x_t2, C * HW, HW, flow, xs, ys,
                           (float*)out, B, C, H, W, layout, st);
  return fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)x_t1, (const __nv_bfloat16*)x_t2, C * HW, HW, flow,
                                 xs, ys, (__nv_bfloat16*)out, B, C, H, W, layout, st);
}

int smow_warp_stack_fwd(const void* x, const float* flow, const float* xs, const float* ys, void* out,
                        int B, int C, int H, int W, int dtype, int layout, void* stream) {
  if (!ys) return fail(SMOW_EINVAL, "null pointer argument");
  if (int e = check_common(x, flow, out, xs, B, C, H, W, dtype, layout)) return e;
  const int64_t HW = (int64_t)H * W;
  cudaStream_t st = (cudaStream_t)stream;
  // stacked NCDHW: frame 1 starts HW elements after frame 0; NDHWC: frame stride is HW*C
  if (dtype == SMOW_F32) {
    const float* p = (const float*)x;
    const float* p2 = p + (layout == SMOW_NCDHW ? HW : HW * C);
    return fwd_impl<float>(p, p2, 2 * C * HW, 2 * HW, flow, xs, ys, (float*)out, B, C, H, W, layout, st);
  }
  const __nv_bfloat16* p = (const __nv_bfloat16*)x;
  const __nv_bfloat16* p2 = p + (layout == SMOW_NCDHW ? HW : HW * C);
  return fwd_impl<__nv_bfloat16>(p, p2, 2 * C * HW, 2 * HW, flow, xs, ys, (__nv_bfloat16*)out, B, C, H, W,
                                 layout, st);
}

int smow_warp_pair_bwd(const void* gout, const void* x_t1, const void* x_t2, const float* flow,
                       const float* xs, const float* ys, void* gx_t1, void* gx_t2, float* gflow, int B,
                       int C, int H, int W, int dtype, int layout, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  if (!x_t2 || !gx_t2 || !gflow || !xs || !ys || !flow) return fail(SMOW_EINVAL, "null pointer argument");
  if (int e = check_common(gout, x_t1, gx_t1, gflow, B, C, H, W, dtype, layout)) return e;
  const int64_t HW = (int64_t)H * W;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMOW_F32)
    return bwd_impl<float>((const float*)gout, (const float*)x_t1, (const float*)x_t2, C * HW, HW, flow,
                           xs, ys, (float*)gx_t1, (float*)gx_t2, gflow, B, C, H, W, layout, workspace, workspace_bytes, st);
  return bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)gout, (const __nv_bfloat16*)x_t1,
                                 (const __nv_bfloat16*)x_t2, C * HW, HW, flow, xs, ys,
                                 (__nv_bfloat16*)gx_t1, (__nv_bfloat16*)gx_t2, gflow, B, C, H, W, layout, workspace,
                                 workspace_bytes, st);
}

int smow_warp_stack_bwd(const void* gout, const void* x, const float* flow, const float* xs,
                        const float* ys, void* gx, float* gflow, int B, int C, int H, int W, int dtype,
                        int layout, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!gflow || !xs || !ys || !flow) return fail(SMOW_EINVAL, "null pointer argument");
  if (int e = check_common(gout, x, gx, gflow, B, C, H, W, dtype, layout)) return e;
  const int64_t HW = (int64_t)H * W;
  const int64_t fs = (layout == SMOW_NCDHW ? HW : HW * C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMOW_F32) {
    const float* p = (const float*)x;
    float* q = (float*)gx;
    return bwd_impl<float>((const float*)gout, p, p + fs, 2 * C * HW, 2 * HW, flow, xs, ys, q, q + fs, gflow,
                           B, C, H, W, layout, workspace, workspace_bytes, st);
  }
  const __nv_bfloat16* p = (const __nv_bfloat16*)x;
  __nv_bfloat16* q = (__nv_bfloat16*)gx;
  return bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)gout, p, p + fs, 2 * C * HW, 2 * HW, flow, xs, ys, q,
                                 q + fs, gflow, B, C, H, W, layout, workspace, workspace_bytes, st);
}

}  // extern "C"
