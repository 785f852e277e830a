// A1, channels-last (NDHWC / channels_last_3d) variants: memory order (B, T, H, W, C).
//
// This is the layout the hot path wants: the C channels of a pixel are contiguous, so each of the four
// bilinear taps is one coalesced 16-byte-vector gather per 4 channels (8 for bf16) straight from global
// memory / L1 — no transposition, no staging pass — and the warped + un-warped slots leave as 16-byte
// stores.  A thread owns (pixel, channel-vector); the q = C/V threads of a pixel sit in adjacent lanes.
//
// Backward (fp32): gx is first set to the un-warped slots, then every (pixel, channel-vector) adds its four
// weighted gradients with ONE vector reduction each (red.global.add.v4.f32, SASS REDG.E.ADD.F32x4 — 4x fewer
// L2 atomics than ATen's scalar scatter), and the flow-gradient partial sums of the q lanes of a pixel are
// combined with warp shuffles.  Any displacement is handled uniformly (no tiles, no far pass).
#include <type_traits>
#include "warp_stack_tiled.cuh"

namespace smow {

template <typename T> struct CVec;
template <> struct CVec<float> { static constexpr int N = 4; };
template <> struct CVec<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> struct Pack { float f[CVec<T>::N]; };

template <typename T> __device__ __forceinline__ Pack<T> ld_pack(const T* p);
template <> __device__ __forceinline__ Pack<float> ld_pack<float>(const float* p) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  Pack<float> r; r.f[0] = v.x; r.f[1] = v.y; r.f[2] = v.z; r.f[3] = v.w;
  return r;
}
template <> __device__ __forceinline__ Pack<__nv_bfloat16> ld_pack<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  Pack<__nv_bfloat16> r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.f[2 * i] = __uint_as_float(w[i] << 16);             // bf16 -> fp32 is a 16-bit shift
    r.f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
template <typename T> __device__ __forceinline__ void st_pack(T* p, const Pack<T>& v);
template <> __device__ __forceinline__ void st_pack<float>(float* p, const Pack<float>& v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.f[0], v.f[1], v.f[2], v.f[3]);
}
template <> __device__ __forceinline__ void st_pack<__nv_bfloat16>(__nv_bfloat16* p, const Pack<__nv_bfloat16>& v) {
  uint4 o;
  uint32_t* w = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.f[2 * i], v.f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = o;
}

// item index -> (pixel, channel-vector); q is a power of two in every SMOW-Net shape
__device__ __forceinline__ void split_item(int idx, int q, int qshift, int& p, int& v) {
  if (qshift >= 0) { p = idx >> qshift; v = idx & (q - 1); }
  else { p = idx / q; v = idx - p * q; }
}

// grid: x = ceil(HW*q / 256), y = 2*B.  x1/x2: frame pointers (element (b,p,c) at b*sB + p*C + c).
template <typename T>
__global__ void __launch_bounds__(256)
warp_fwd_ndhwc_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int64_t sB, const float* __restrict__ flow,
                      const float* __restrict__ xs, const float* __restrict__ ys, T* __restrict__ out, int C, int H,
                      int W, int q, int qshift) {
  constexpr int V = CVec<T>::N;
  const int HW = H * W;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= HW * q) return;
  int p, v;
  split_item(idx, q, qshift, p, v);
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  const int h = p / W, w = p - h * W;
  const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
  const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
  const float nw = __fmul_rn(fp.wx0, fp.wy0), ne = __fmul_rn(fp.wx1, fp.wy0);
  const float sw = __fmul_rn(fp.wx0, fp.wy1), se = __fmul_rn(fp.wx1, fp.wy1);
  const T* src = (t ? x2 : x1) + b * sB + v * V;
  const int o_nw = fp.y0 * W + fp.x0;
  const Pack<T> a = ld_pack<T>(src + (int64_t)o_nw * C);
  Pack<T> r;
#pragma unroll
  for (int j = 0; j < V; ++j) r.f[j] = __fmul_rn(a.f[j], nw);          // ATen order nw, ne, sw, se; skip OOB taps
  if (fp.x1ok) {
    const Pack<T> c = ld_pack<T>(src + (int64_t)(o_nw + 1) * C);
#pragma unroll
    for (int j = 0; j < V; ++j) r.f[j] = fmaf(c.f[j], ne, r.f[j]);
  }
  if (fp.y1ok) {
    const Pack<T> c = ld_pack<T>(src + (int64_t)(o_nw + W) * C);
#pragma unroll
    for (int j = 0; j < V; ++j) r.f[j] = fmaf(c.f[j], sw, r.f[j]);
  }
  if (fp.x1ok && fp.y1ok) {
    const Pack<T> c = ld_pack<T>(src + (int64_t)(o_nw + W + 1) * C);
#pragma unroll
    for (int j = 0; j < V; ++j) r.f[j] = fmaf(c.f[j], se, r.f[j]);
  }
  T* ob = out + ((int64_t)b * 4 * HW + p) * C + v * V;      // slot s at + s*HW*C
  st_pack<T>(ob + (int64_t)(1 + t) * HW * C, r);
  *reinterpret_cast<uint4*>(ob + (int64_t)(t ? 3 : 0) * HW * C) =
      __ldg(reinterpret_cast<const uint4*>(src + (int64_t)p * C));    // un-warped slot: bit-exact copy
}

// gx[b,t] = gout[b, pass(t)]  (+ zero the flow gradient when it is accumulated with atomics)
template <typename T>
__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_init_kernel(const T* __restrict__ gout, T* __restrict__ gx1, T* __restrict__ gx2, int64_t sB,
                           float* __restrict__ gflow, int C, int HW, int q, bool zero_gflow, const int* __restrict__ hdr) {
  constexpr int V = CVec<T>::N;
  if (hdr != nullptr) {               // gather mode: only the (rare, wide-C) atomic flow-gradient needs zeroing
    const int nx = hdr[1] - hdr[0] + 2, ny = hdr[3] - hdr[2] + 2;
    if (nx > 0 && ny > 0 && nx * ny <= 49) {
      const int i = blockIdx.x * 256 + threadIdx.x;
      if (zero_gflow && i < HW) {
        const int64_t fo = ((int64_t)((blockIdx.y >> 1) * 2) * 2 + (blockIdx.y & 1)) * HW + i;
        gflow[fo] = 0.f; gflow[fo + 2 * (int64_t)HW] = 0.f;
      }
      return;
    }
  }
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= HW * q) return;
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  const int64_t e = (int64_t)idx * V;                        // element offset inside one (b, frame) plane
  *reinterpret_cast<uint4*>((t ? gx2 : gx1) + b * sB + e) =
      __ldg(reinterpret_cast<const uint4*>(gout + ((int64_t)(b * 4 + (t ? 3 : 0)) * HW) * C + e));
  if (zero_gflow && idx < HW) {
    const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + idx;
    gflow[fo] = 0.f; gflow[fo + 2 * (int64_t)HW] = 0.f;
  }
}

__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_scatter_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                              int64_t sB, const float* __restrict__ flow, const float* __restrict__ xs,
                              const float* __restrict__ ys, float* __restrict__ gx1, float* __restrict__ gx2,
                              float* __restrict__ gflow, int C, int H, int W, int q, int qshift, bool shuffle_reduce,
                              const int* __restrict__ hdr) {
  if (hdr != nullptr) {               // the gather kernel already produced this launch's gradients
    const int nx = hdr[1] - hdr[0] + 2, ny = hdr[3] - hdr[2] + 2;
    if (nx > 0 && ny > 0 && nx * ny <= 49) return;
  }
  const int HW = H * W;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool live = idx < HW * q;
  int p = 0, v = 0;
  if (live) split_item(idx, q, qshift, p, v);
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  float gix = 0.f, giy = 0.f, mx = 0.f, my = 0.f;
  int64_t fo = 0;
  if (live) {
    const int h = p / W, w = p - h * W;
    fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
    const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
    const float nw = __fmul_rn(fp.wx0, fp.wy0), ne = __fmul_rn(fp.wx1, fp.wy0);
    const float sw = __fmul_rn(fp.wx0, fp.wy1), se = __fmul_rn(fp.wx1, fp.wy1);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gout + ((int64_t)(b * 4 + 1 + t) * HW + p) * C + v * 4));
    const int o_nw = fp.y0 * W + fp.x0;
    const float* src = (t ? x2 : x1) + b * sB + v * 4;
    float* dst = (t ? gx2 : gx1) + b * sB + v * 4;
    auto tap = [&](int off, float wgt, float dgx, float dgy) {
      atomicAdd(reinterpret_cast<float4*>(dst + (int64_t)off * C), make_float4(wgt * g.x, wgt * g.y, wgt * g.z, wgt * g.w));
      const float4 xv = __ldg(reinterpret_cast<const float4*>(src + (int64_t)off * C));
      const float dot = fmaf(xv.w, g.w, fmaf(xv.z, g.z, fmaf(xv.y, g.y, __fmul_rn(xv.x, g.x))));
      gix = fmaf(dgx, dot, gix);
      giy = fmaf(dgy, dot, giy);
    };
    tap(o_nw, nw, -fp.wy0, -fp.wx0);
    if (fp.x1ok) tap(o_nw + 1, ne, fp.wy0, -fp.wx1);
    if (fp.y1ok) tap(o_nw + W, sw, -fp.wy1, fp.wx0);
    if (fp.x1ok && fp.y1ok) tap(o_nw + W + 1, se, fp.wy1, fp.wx1);
    mx = __fmul_rn(fp.gx_gate, __fmul_rn((float)(W - 1), 0.5f));
    my = __fmul_rn(fp.gy_gate, __fmul_rn((float)(H - 1), 0.5f));
  }
  if (shuffle_reduce) {      // q is a power of two: the lanes of a pixel are adjacent, butterfly over min(q, 32) of them
    const int span = q < 32 ? q : 32;
    for (int d = span >> 1; d > 0; d >>= 1) {
      gix += __shfl_xor_sync(0xffffffffu, gix, d);
      giy += __shfl_xor_sync(0xffffffffu, giy, d);
    }
    if (live && (v & (span - 1)) == 0) {
      if (q <= 32) {
        gflow[fo] = __fdiv_rn(__fmul_rn(mx, gix), (float)W);
        gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, giy), (float)H);
      } else {               // several warps share a pixel: one atomic per warp into the zeroed gradient
        atomicAdd(gflow + fo, __fdiv_rn(__fmul_rn(mx, gix), (float)W));
        atomicAdd(gflow + fo + 2 * (int64_t)HW, __fdiv_rn(__fmul_rn(my, giy), (float)H));
      }
    }
  } else if (live) {
    atomicAdd(gflow + fo, __fdiv_rn(__fmul_rn(mx, gix), (float)W));
    atomicAdd(gflow + fo + 2 * (int64_t)HW, __fdiv_rn(__fmul_rn(my, giy), (float)H));
  }
}


// ------------------------------------------------------------------------------
// deterministic gather backward (needs the caller's workspace)
// ------------------------------------------------------------------------------
// workspace: [0..3] int header {dxmin, dxmax, dymin, dymax} = range of the integer displacement
// (anchor - pixel) over the whole launch, then two float planes (ix, iy) of B*2*HW clipped sample coordinates.
constexpr int GATHER_MAX_WINDOW = 49;     // candidates per target beyond which the atomic scatter takes over

__device__ __forceinline__ bool gather_enabled(const int* hdr) {
  const int nx = hdr[1] - hdr[0] + 2, ny = hdr[3] - hdr[2] + 2;
  return nx > 0 && ny > 0 && nx * ny <= GATHER_MAX_WINDOW;
}

__global__ void warp_bwd_ndhwc_hdr_kernel(int* hdr) {
  if (threadIdx.x < 4) hdr[threadIdx.x] = (threadIdx.x & 1) ? -(1 << 30) : (1 << 30);
}

__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_stat_kernel(const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                           int* __restrict__ hdr, float* __restrict__ cix, float* __restrict__ ciy, int H, int W) {
  __shared__ int red[4];
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  if (threadIdx.x < 4) red[threadIdx.x] = (threadIdx.x & 1) ? -(1 << 30) : (1 << 30);
  __syncthreads();
  int dx0 = 1 << 30, dx1 = -(1 << 30), dy0 = 1 << 30, dy1 = -(1 << 30);
  if (p < HW) {
    const int h = p / W, w = p - h * W;
    const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
    const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
    const int64_t o = (int64_t)(b * 2 + t) * HW + p;
    cix[o] = __fadd_rn((float)fp.x0, fp.wx1);          // = ix exactly (wx1 = ix - floor(ix))
    ciy[o] = __fadd_rn((float)fp.y0, fp.wy1);
    dx0 = dx1 = fp.x0 - w; dy0 = dy1 = fp.y0 - h;
  }
  dx0 = __reduce_min_sync(0xffffffffu, dx0); dx1 = __reduce_max_sync(0xffffffffu, dx1);
  dy0 = __reduce_min_sync(0xffffffffu, dy0); dy1 = __reduce_max_sync(0xffffffffu, dy1);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(red + 0, dx0); atomicMax(red + 1, dx1); atomicMin(red + 2, dy0); atomicMax(red + 3, dy1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMin(hdr + 0, red[0]); atomicMax(hdr + 1, red[1]); atomicMin(hdr + 2, red[2]); atomicMax(hdr + 3, red[3]);
  }
}

// thread = (pixel, 4-channel vector).  Target side: gx = gout[pass] + sum over the sources whose bilinear
// footprint covers this pixel (found by scanning the coordinate planes in a fixed order: bit-reproducible);
// source side: flow-gradient sums of this pixel's own footprint, combined over the lanes of the pixel.
__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_gather_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                             int64_t sB, const float* __restrict__ flow, const float* __restrict__ xs,
                             const float* __restrict__ ys, const int* __restrict__ hdr, const float* __restrict__ cix,
                             const float* __restrict__ ciy, float* __restrict__ gx1, float* __restrict__ gx2,
                             float* __restrict__ gflow, int C, int H, int W, int q, int qshift) {
  if (!gather_enabled(hdr)) return;                       // the atomic scatter kernels handle this launch
  const int dxlo = hdr[0], dxhi = hdr[1], dylo = hdr[2], dyhi = hdr[3];
  const int HW = H * W;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool live = idx < HW * q;
  int p = 0, v = 0;
  if (live) split_item(idx, q, qshift, p, v);
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  float gix = 0.f, giy = 0.f, mx = 0.f, my = 0.f;
  int64_t fo = 0;
  if (live) {
    const int h = p / W, w = p - h * W;
    fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
    const float* gw = gout + ((int64_t)(b * 4 + 1 + t) * HW) * C + v * 4;      // warped-slot gradient plane
    // ---- target side ----
    float4 acc = __ldg(reinterpret_cast<const float4*>(gout + ((int64_t)(b * 4 + (t ? 3 : 0)) * HW + p) * C + v * 4));
    const float* px = cix + (int64_t)(b * 2 + t) * HW;
    const float* py = ciy + (int64_t)(b * 2 + t) * HW;
    const int sy_a = max(0, h - 1 - dyhi), sy_b = min(H - 1, h - dylo);
    const int sx_a = max(0, w - 1 - dxhi), sx_b = min(W - 1, w - dxlo);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sy = sy_a; sy <= sy_b; ++sy) {
      for (int sx = sx_a; sx <= sx_b; ++sx) {
        const int s = sy * W + sx;
        const float ix = __ldg(px + s), iy = __ldg(py + s);
        const float x0f = floorf(ix), y0f = floorf(iy);
        const unsigned ex = (unsigned)(w - (int)x0f), ey = (unsigned)(h - (int)y0f);
        if ((ex | ey) > 1u) continue;
        const float wx = ex ? __fsub_rn(ix, x0f) : __fsub_rn(__fadd_rn(x0f, 1.f), ix);
        const float wy = ey ? __fsub_rn(iy, y0f) : __fsub_rn(__fadd_rn(y0f, 1.f), iy);
        const float wgt = __fmul_rn(wx, wy);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gw + (int64_t)s * C));
        sum.x = fmaf(wgt, g.x, sum.x); sum.y = fmaf(wgt, g.y, sum.y);
        sum.z = fmaf(wgt, g.z, sum.z); sum.w = fmaf(wgt, g.w, sum.w);
      }
    }
    acc.x = __fadd_rn(acc.x, sum.x); acc.y = __fadd_rn(acc.y, sum.y);
    acc.z = __fadd_rn(acc.z, sum.z); acc.w = __fadd_rn(acc.w, sum.w);
    *reinterpret_cast<float4*>((t ? gx2 : gx1) + b * sB + (int64_t)p * C + v * 4) = acc;
    // ---- source side ----
    const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gw + (int64_t)p * C));
    const int o_nw = fp.y0 * W + fp.x0;
    const float* src = (t ? x2 : x1) + b * sB + v * 4;
    auto tap = [&](int off, float dgx, float dgy) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(src + (int64_t)off * C));
      const float dot = fmaf(xv.w, g.w, fmaf(xv.z, g.z, fmaf(xv.y, g.y, __fmul_rn(xv.x, g.x))));
      gix = fmaf(dgx, dot, gix);
      giy = fmaf(dgy, dot, giy);
    };
    tap(o_nw, -fp.wy0, -fp.wx0);
    if (fp.x1ok) tap(o_nw + 1, fp.wy0, -fp.wx1);
    if (fp.y1ok) tap(o_nw + W, -fp.wy1, fp.wx0);
    if (fp.x1ok && fp.y1ok) tap(o_nw + W + 1, fp.wy1, fp.wx1);
    mx = __fmul_rn(fp.gx_gate, __fmul_rn((float)(W - 1), 0.5f));
    my = __fmul_rn(fp.gy_gate, __fmul_rn((float)(H - 1), 0.5f));
  }
  const int span = q < 32 ? q : 32;       // q is a power of two here (checked by the host)
  for (int d = span >> 1; d > 0; d >>= 1) {
    gix += __shfl_xor_sync(0xffffffffu, gix, d);
    giy += __shfl_xor_sync(0xffffffffu, giy, d);
  }
  if (live && (v & (span - 1)) == 0) {
    if (q <= 32) {
      gflow[fo] = __fdiv_rn(__fmul_rn(mx, gix), (float)W);
      gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, giy), (float)H);
    } else {
      atomicAdd(gflow + fo, __fdiv_rn(__fmul_rn(mx, gix), (float)W));
      atomicAdd(gflow + fo + 2 * (int64_t)HW, __fdiv_rn(__fmul_rn(my, giy), (float)H));
    }
  }
}

static int ilog2_exact(int q) {
  for (int s = 0; s < 31; ++s)
    if ((1 << s) == q) return s;
  return -1;
}

template <typename T>
int warp_fwd_ndhwc(const T* x1, const T* x2, int64_t sB, const float* flow, const float* xs, const float* ys, T* out,
                   int B, int C, int H, int W, cudaStream_t st) {
  constexpr int V = CVec<T>::N;
  if (C % V || !aligned16(x1) || !aligned16(x2) || !aligned16(out) || sB % V)
    return fail(SMOW_EALIGN, "NDHWC warp needs C %% %d == 0 and 16 B aligned tensors", V);
  const int q = C / V;
  if ((int64_t)H * W * q >= (1ll << 31)) return fail(SMOW_ERANGE, "plane too large");
  dim3 grid((unsigned)(((int64_t)H * W * q + 255) / 256), 2 * B);
  warp_fwd_ndhwc_kernel<T><<<grid, 256, 0, st>>>(x1, x2, sB, flow, xs, ys, out, C, H, W, q, ilog2_exact(q));
  count_launch();
  return check_launch("warp_fwd_ndhwc");
}

template <typename T>
int warp_bwd_ndhwc(const T* gout, const T* x1, const T* x2, int64_t sB, const float* flow, const float* xs,
                   const float* ys, T* gx1, T* gx2, float* gflow, int B, int C, int H, int W, void* ws,
                   int64_t ws_bytes, cudaStream_t st) {
  if constexpr (!std::is_same<T, float>::value) {
    return fail(SMOW_EDTYPE, "NDHWC warp backward is built for fp32 only (use the NCDHW layout for bf16)");
  } else {
    if (C % 4 || !aligned16(gout) || !aligned16(x1) || !aligned16(x2) || !aligned16(gx1) || !aligned16(gx2) || sB % 4)
      return fail(SMOW_EALIGN, "NDHWC warp needs C %% 4 == 0 and 16 B aligned tensors");
    const int q = C / 4, qs = ilog2_exact(q);
    if ((int64_t)H * W * q >= (1ll << 31)) return fail(SMOW_ERANGE, "plane too large");
    const bool shuffle = qs >= 0;
    const int HW = H * W;
    dim3 grid((unsigned)(((int64_t)HW * q + 255) / 256), 2 * B);
    // deterministic gather when the caller lent a workspace (and the lanes of a pixel can be shuffle-reduced)
    const int64_t need = 64 + (int64_t)B * 2 * 2 * HW * (int64_t)sizeof(float);
    int* hdr = nullptr;
    int launches = 2;
    if (ws != nullptr && ws_bytes >= need && shuffle && q <= 32 && aligned16(ws) && option(OPT_WARP_BWD_VARIANT) != 0) {
      hdr = reinterpret_cast<int*>(ws);
      float* cix = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 64);
      float* ciy = cix + (int64_t)B * 2 * HW;
      warp_bwd_ndhwc_hdr_kernel<<<1, 32, 0, st>>>(hdr);   // min/max sentinels
      warp_bwd_ndhwc_stat_kernel<<<dim3((HW + 255) / 256, 2 * B), 256, 0, st>>>(flow, xs, ys, hdr, cix, ciy, H, W);
      warp_bwd_ndhwc_gather_kernel<<<grid, 256, 0, st>>>(gout, x1, x2, sB, flow, xs, ys, hdr, cix, ciy, gx1, gx2, gflow,
                                                         C, H, W, q, qs);
      launches += 3;
    }
    warp_bwd_ndhwc_init_kernel<float><<<grid, 256, 0, st>>>(gout, gx1, gx2, sB, gflow, C, HW, q, !shuffle || q > 32, hdr);
    warp_bwd_ndhwc_scatter_kernel<<<grid, 256, 0, st>>>(gout, x1, x2, sB, flow, xs, ys, gx1, gx2, gflow, C, H, W, q, qs,
                                                        shuffle, hdr);
    count_launch(launches);
    return check_launch("warp_bwd_ndhwc");
  }
}

#define INST(T)                                                                                              \
  template int warp_fwd_ndhwc<T>(const T*, const T*, int64_t, const float*, const float*, const float*, T*, \
                                 int, int, int, int, cudaStream_t);                                         \
  template int warp_bwd_ndhwc<T>(const T*, const T*, const T*, int64_t, const float*, const float*,         \
                                 const float*, T*, T*, float*, int, int, int, int, void*, int64_t, cudaStream_t);
INST(float)
INST(__nv_bfloat16)
}  // namespace smow
