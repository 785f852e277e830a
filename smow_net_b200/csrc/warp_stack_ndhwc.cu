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
#include <atomic>
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

// Same forward with the coordinate chain de-duplicated across the lanes of a pixel (q = C/V a power of two <= 32):
// a warp owns 32 consecutive pixels.  Lane l runs the reference's coordinate chain for pixel l ONCE (~100 instructions
// per 32 pixels instead of per (pixel, vector)); then, in q steps of 32/q pixels, every lane fetches the footprint of
// the pixel it serves with indexed shuffles and does only the gather / FMA / store work.  Arithmetic and rounding are
// unchanged (bit-identical results); 3.5x fewer instructions at C = 32.
// grid: x = ceil(HW / 256) (8 warps x 32 pixels), y = 2*B.
template <typename T>
__device__ __forceinline__ void warp_fwd_ndhwc_shfl_body(const T* __restrict__ x1, const T* __restrict__ x2, int64_t sB,
                           const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                           T* __restrict__ out, int C, int H, int W, int q, int qshift, int wshift) {
  constexpr int V = CVec<T>::N;
  const int HW = H * W;
  const int lane = threadIdx.x & 31;
  const int pbase = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32;
  if (pbase >= HW) return;                                       // warp-uniform
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  // ---- one coordinate chain per pixel ----
  int own_o = 0, own_flags = 0;
  float own_nw = 0.f, own_ne = 0.f, own_sw = 0.f, own_se = 0.f;
  {
    const int p = pbase + lane;
    if (p < HW) {
      const int h = wshift >= 0 ? (p >> wshift) : (p / W), w = p - h * W;
      const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
      const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
      own_nw = __fmul_rn(fp.wx0, fp.wy0); own_ne = __fmul_rn(fp.wx1, fp.wy0);
      own_sw = __fmul_rn(fp.wx0, fp.wy1); own_se = __fmul_rn(fp.wx1, fp.wy1);
      own_o = fp.y0 * W + fp.x0;
      own_flags = (fp.x1ok ? 1 : 0) | (fp.y1ok ? 2 : 0);
    }
  }
  // ---- q steps: lanes = (32/q pixels) x (q channel vectors) ----
  // More than 32 vectors per pixel (fp32 C = 256 / 512): blockIdx.z walks groups of 32 vectors, q = 32 per group (the chain
  // is then run once per group: ~100 instructions per 32 pixels next to 32 x 32 gathers)
  const int v = (lane & (q - 1)) + 32 * blockIdx.z, sub = lane >> qshift, ppw = 32 >> qshift;
  const T* src = (t ? x2 : x1) + b * sB + v * V;
  T* ob = out + (int64_t)b * 4 * HW * C + v * V;
  const int64_t slotC = (int64_t)HW * C;
  if constexpr (std::is_same<T, float>::value) {
    // fp32: 48 registers / 5 CTAs per SM beat the unrolled form below (60 registers): 0.83 vs 0.76 at C = 64
    for (int j = 0; j < q; ++j) {
      const int sl = j * ppw + sub;                                 // lane that holds this pixel's footprint
      const int o_nw = __shfl_sync(0xffffffffu, own_o, sl), flags = __shfl_sync(0xffffffffu, own_flags, sl);
      const float nw = __shfl_sync(0xffffffffu, own_nw, sl), ne = __shfl_sync(0xffffffffu, own_ne, sl);
      const float sw = __shfl_sync(0xffffffffu, own_sw, sl), se = __shfl_sync(0xffffffffu, own_se, sl);
      const int p = pbase + sl;
      if (p >= HW) continue;
      const T* tp = src + (int64_t)o_nw * C;
      const Pack<T> a = ld_pack<T>(tp);
      const uint4 pass = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)p * C));
      Pack<T> r;
  #pragma unroll
      for (int k = 0; k < V; ++k) r.f[k] = __fmul_rn(a.f[k], nw);        // ATen order nw, ne, sw, se; skip OOB taps
      if (flags & 1) {
        const Pack<T> c = ld_pack<T>(tp + C);
  #pragma unroll
        for (int k = 0; k < V; ++k) r.f[k] = fmaf(c.f[k], ne, r.f[k]);
      }
      if (flags & 2) {
        const Pack<T> c = ld_pack<T>(tp + (int64_t)W * C);
  #pragma unroll
        for (int k = 0; k < V; ++k) r.f[k] = fmaf(c.f[k], sw, r.f[k]);
      }
      if (flags == 3) {
        const Pack<T> c = ld_pack<T>(tp + (int64_t)(W + 1) * C);
  #pragma unroll
        for (int k = 0; k < V; ++k) r.f[k] = fmaf(c.f[k], se, r.f[k]);
      }
      T* o = ob + (int64_t)p * C;
      st_pack<T>(o + (1 + t) * slotC, r);
      *reinterpret_cast<uint4*>(o + (t ? 3 : 0) * slotC) = pass;            // un-warped slot: bit-exact copy
    }
  } else {
    // branch-free body (out-of-image taps are re-pointed at the nw tap and dropped from the sum by a select: same bits as the
    // skipped tap), so that the unrolled loop keeps the 5 loads of several pixels in flight: the kernel waits on its gathers
    // (ncu: 22 long-scoreboard stalls per issued instruction at 57 % occupancy)
  #pragma unroll 4
    for (int j = 0; j < q; ++j) {
      const int sl = j * ppw + sub;                                 // lane that holds this pixel's footprint
      const int o_nw = __shfl_sync(0xffffffffu, own_o, sl), flags = __shfl_sync(0xffffffffu, own_flags, sl);
      const float nw = __shfl_sync(0xffffffffu, own_nw, sl), ne = __shfl_sync(0xffffffffu, own_ne, sl);
      const float sw = __shfl_sync(0xffffffffu, own_sw, sl), se = __shfl_sync(0xffffffffu, own_se, sl);
      const int p = pbase + sl;
      const bool live = p < HW;                                     // dead lanes carry o_nw = 0, flags = 0: loads stay in bounds
      const bool f1 = flags & 1, f2 = flags & 2, f3 = flags == 3;
      const T* tp = src + (int64_t)o_nw * C;
      const Pack<T> a = ld_pack<T>(tp);
      const Pack<T> c1 = ld_pack<T>(tp + (f1 ? C : 0));
      const Pack<T> c2 = ld_pack<T>(tp + (f2 ? (int64_t)W * C : 0));
      const Pack<T> c3 = ld_pack<T>(tp + (f3 ? (int64_t)(W + 1) * C : 0));
      const uint4 pass = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)(live ? p : pbase) * C));
      Pack<T> r;
  #pragma unroll
      for (int k = 0; k < V; ++k) {                                 // ATen order nw, ne, sw, se; out-of-image taps skipped
        float t0 = __fmul_rn(a.f[k], nw);
        t0 = f1 ? fmaf(c1.f[k], ne, t0) : t0;
        t0 = f2 ? fmaf(c2.f[k], sw, t0) : t0;
        t0 = f3 ? fmaf(c3.f[k], se, t0) : t0;
        r.f[k] = t0;
      }
      if (live) {
        T* o = ob + (int64_t)p * C;
        st_pack<T>(o + (1 + t) * slotC, r);
        *reinterpret_cast<uint4*>(o + (t ? 3 : 0) * slotC) = pass;          // un-warped slot: bit-exact copy
      }
    }
  }
}

// Two launch shapes of the same body (measured on B200, benchmarks/fwd_cold_probe.py): fp32 is best left to the compiler
// (48 registers, 5 CTAs per SM: 0.83-0.89; forcing 5 CTAs gives 46 registers and 0.77-0.84); the unrolled bf16 loop wants 4 CTAs
// per SM (64 registers, 32 B of spills: 0.71-0.77 instead of 0.65-0.75 at 69 registers / 3 CTAs).
template <typename T>
__global__ void __launch_bounds__(256)
warp_fwd_ndhwc_shfl_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int64_t sB,
                           const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                           T* __restrict__ out, int C, int H, int W, int q, int qshift, int wshift) {
  warp_fwd_ndhwc_shfl_body<T>(x1, x2, sB, flow, xs, ys, out, C, H, W, q, qshift, wshift);
}
template <typename T>
__global__ void __launch_bounds__(256, 4)
warp_fwd_ndhwc_shfl4_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int64_t sB,
                            const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                            T* __restrict__ out, int C, int H, int W, int q, int qshift, int wshift) {
  warp_fwd_ndhwc_shfl_body<T>(x1, x2, sB, flow, xs, ys, out, C, H, W, q, qshift, wshift);
}

// ------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------
// Two strategies behind one entry point:
//  (G) deterministic gather — needs the caller's workspace.  stat: clipped sample coordinates + packed anchors
//      of every pixel, and the launch-wide range of integer displacements; list: ONE thread per pixel scans
//      the few sources that can cover it and writes a (weight, source) list — per pixel, not per channel;
//      apply: one thread per (pixel, 4-channel vector) reads the list (broadcast loads) and does
//      gx = gout[pass] + sum w * gout[warp][src] with plain 16-byte stores, plus the flow-gradient sums.
//      No float atomics, no zero-fill, fixed summation order => bit-reproducible.
//  (S) vector-atomic scatter — any displacement: gx <- gout[pass], then one red.global.add.v4.f32 per tap.
// The choice is made ON THE DEVICE (no host sync): stat/list publish the displacement window and a list
// overflow flag in the workspace header; (G)'s apply kernel and (S)'s kernels each read it and exactly one of
// the two families does the work, the other exits at once.
constexpr int GATHER_MAX_WINDOW = 49;   // candidate sources per target beyond which (S) takes over
constexpr int LIST_K = 9;               // list entries per target (a 3x3 window can never overflow it); more => (S)

struct BwdWs {        // views into the caller's workspace
  int* hdr;           // {dxmin, dxmax, dymin, dymax, overflow}
  float* cix; float* ciy; int* can;     // per pixel-frame: clipped coordinates, packed anchor
  int* cnt; float* lw; int* ls;         // per pixel-frame: list length; [LIST_K][N] weights and source pixels
};
__host__ __device__ inline int64_t bwd_ws_bytes(int64_t n) { return 64 + n * (3 + 1 + 2 * LIST_K) * 4; }
static BwdWs carve_ws(void* ws, int64_t n) {
  BwdWs w;
  w.hdr = reinterpret_cast<int*>(ws);
  w.cix = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 64);
  w.ciy = w.cix + n;
  w.can = reinterpret_cast<int*>(w.ciy + n);
  w.cnt = w.can + n;
  w.lw = reinterpret_cast<float*>(w.cnt + n);
  w.ls = reinterpret_cast<int*>(w.lw + LIST_K * n);
  return w;
}
__device__ __forceinline__ bool gather_active(const int* hdr) {
  const int nx = hdr[1] - hdr[0] + 2, ny = hdr[3] - hdr[2] + 2;
  return nx > 0 && ny > 0 && nx * ny <= GATHER_MAX_WINDOW && hdr[4] == 0;
}

__global__ void warp_bwd_ndhwc_hdr_kernel(int* hdr) {
  if (threadIdx.x < 4) hdr[threadIdx.x] = (threadIdx.x & 1) ? -(1 << 30) : (1 << 30);
  if (threadIdx.x == 4) hdr[4] = 0;
}

__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_stat_kernel(const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                           BwdWs ws, int H, int W) {
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
    ws.cix[o] = __fadd_rn((float)fp.x0, fp.wx1);          // = ix exactly (wx1 = ix - floor(ix))
    ws.ciy[o] = __fadd_rn((float)fp.y0, fp.wy1);
    ws.can[o] = (fp.y0 << 16) | fp.x0;
    dx0 = dx1 = fp.x0 - w; dy0 = dy1 = fp.y0 - h;
  }
  dx0 = __reduce_min_sync(0xffffffffu, dx0); dx1 = __reduce_max_sync(0xffffffffu, dx1);
  dy0 = __reduce_min_sync(0xffffffffu, dy0); dy1 = __reduce_max_sync(0xffffffffu, dy1);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(red + 0, dx0); atomicMax(red + 1, dx1); atomicMin(red + 2, dy0); atomicMax(red + 3, dy1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // same-address L2 atomics serialise: only blocks that widen the range issue one
    if (red[0] < *(volatile int*)(ws.hdr + 0)) atomicMin(ws.hdr + 0, red[0]);
    if (red[1] > *(volatile int*)(ws.hdr + 1)) atomicMax(ws.hdr + 1, red[1]);
    if (red[2] < *(volatile int*)(ws.hdr + 2)) atomicMin(ws.hdr + 2, red[2]);
    if (red[3] > *(volatile int*)(ws.hdr + 3)) atomicMax(ws.hdr + 3, red[3]);
  }
}

// one thread per target pixel-frame: scan the candidate window in fixed (row-major) order
__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_list_kernel(BwdWs ws, int H, int W, int64_t n_pf) {
  const int nx = ws.hdr[1] - ws.hdr[0] + 2, ny = ws.hdr[3] - ws.hdr[2] + 2;
  if (!(nx > 0 && ny > 0 && nx * ny <= GATHER_MAX_WINDOW)) return;
  const int dxlo = ws.hdr[0], dxhi = ws.hdr[1], dylo = ws.hdr[2], dyhi = ws.hdr[3];
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int64_t plane = (int64_t)blockIdx.y * HW;          // blockIdx.y = b*2 + t
  const int h = p / W, w = p - h * W;
  const int* pa = ws.can + plane;
  const float* px = ws.cix + plane;
  const float* py = ws.ciy + plane;
  const int sy_a = max(0, h - 1 - dyhi), sy_b = min(H - 1, h - dylo);
  const int sx_a = max(0, w - 1 - dxhi), sx_b = min(W - 1, w - dxlo);
  int n = 0;
  for (int sy = sy_a; sy <= sy_b; ++sy) {
    const int row = sy * W;
    for (int sx = sx_a; sx <= sx_b; ++sx) {
      const int an = __ldg(pa + row + sx);
      const int x0 = an & 0xffff, y0 = an >> 16;
      const unsigned ex = (unsigned)(w - x0), ey = (unsigned)(h - y0);
      if ((ex | ey) > 1u) continue;
      const float ix = __ldg(px + row + sx), iy = __ldg(py + row + sx);
      const float x0f = (float)x0, y0f = (float)y0;
      const float wx = ex ? __fsub_rn(ix, x0f) : __fsub_rn(__fadd_rn(x0f, 1.f), ix);
      const float wy = ey ? __fsub_rn(iy, y0f) : __fsub_rn(__fadd_rn(y0f, 1.f), iy);
      if (n < LIST_K) {
        ws.lw[(int64_t)n * n_pf + plane + p] = __fmul_rn(wx, wy);
        ws.ls[(int64_t)n * n_pf + plane + p] = row + sx;
      }
      ++n;
    }
  }
  ws.cnt[plane + p] = n;
  if (n > LIST_K) ws.hdr[4] = 1;          // benign race: every writer stores 1
}

// one thread per (pixel, 4-channel vector)
__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_apply_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                            int64_t sB, const float* __restrict__ flow, const float* __restrict__ xs,
                            const float* __restrict__ ys, BwdWs ws, float* __restrict__ gx1, float* __restrict__ gx2,
                            float* __restrict__ gflow, int C, int H, int W, int q, int qshift, int64_t n_pf) {
  if (!gather_active(ws.hdr)) return;
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
    const int64_t pf = (int64_t)blockIdx.y * HW + p;
    const float* gw = gout + ((int64_t)(b * 4 + 1 + t) * HW) * C + v * 4;      // warped-slot gradient plane
    // ---- target side: pass-through + gathered scatter ----
    const float4 pass = __ldg(reinterpret_cast<const float4*>(gout + ((int64_t)(b * 4 + (t ? 3 : 0)) * HW + p) * C + v * 4));
    const int n = __ldg(ws.cnt + pf);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < n; ++j) {
      const float wgt = __ldg(ws.lw + (int64_t)j * n_pf + pf);
      const int s = __ldg(ws.ls + (int64_t)j * n_pf + pf);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gw + (int64_t)s * C));
      sum.x = fmaf(wgt, g.x, sum.x); sum.y = fmaf(wgt, g.y, sum.y);
      sum.z = fmaf(wgt, g.z, sum.z); sum.w = fmaf(wgt, g.w, sum.w);
    }
    *reinterpret_cast<float4*>((t ? gx2 : gx1) + b * sB + (int64_t)p * C + v * 4) =
        make_float4(__fadd_rn(pass.x, sum.x), __fadd_rn(pass.y, sum.y), __fadd_rn(pass.z, sum.z), __fadd_rn(pass.w, sum.w));
    // ---- source side: flow-gradient sums of this pixel's own footprint ----
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
  for (int d = q >> 1; d > 0; d >>= 1) {      // q <= 32 and a power of two (host-checked): lanes of a pixel are adjacent
    gix += __shfl_xor_sync(0xffffffffu, gix, d);
    giy += __shfl_xor_sync(0xffffffffu, giy, d);
  }
  if (live && v == 0) {
    gflow[fo] = __fdiv_rn(__fmul_rn(mx, gix), (float)W);
    gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, giy), (float)H);
  }
}

// ---- (T) tile gather: no workspace, no zero-fill, no atomics for displacements under one pixel ------------------
// A CTA owns a tile of R x 32 pixels of one (pair, frame) plane of gx and writes it exactly once.
//  stage  : the warped-slot gradient of the tile plus a one-pixel halo goes to shared memory with 16-byte
//           asynchronous copies (cp.async, SASS LDGSTS: no register staging, the whole tile is in flight at once;
//           halo pixels outside the image are zero-filled), 32 channels at a time;
//  phase 0: meanwhile the clipped sample coordinates (ix, iy) and flow-gradient gates of the (R+2) x 34 sources
//           are computed once per pixel (not per channel vector) into shared memory;
//  phase 1: one thread per target pixel probes its 3x3 sources: source s covers target p with the separable
//           weight wx*wy, wx = (px+1)-ix if ix >= px else ix-(px-1) — the very subtractions ATen's
//           (ix_se-ix)/(ix-ix_nw) perform, so the weights are bit-identical; 9 weights per pixel (0 = no cover);
//  phase 2: one thread per (target pixel, channel vector): gx = gout[pass] + sum_j w_j * staged[j] (9 LDS.128 at
//           immediate offsets, fixed order => bit-reproducible), then the source-side work of the same pixel:
//           4 taps of x against its own staged gradient -> flow-gradient sums, shuffle-reduced over the lanes of
//           the pixel.  Global traffic per item = the forward kernel's (pass + 4 taps in, one vector out).
// Taps farther than one pixel from their source (|tap - source| > 1 on either axis) are not visible to a 3x3
// probe: warp_bwd_ndhwc_far_kernel (stream-ordered after the tile kernel) re-derives exactly that predicate
// from the flow and adds them with vector reductions.  For sub-pixel flows it reads the flow and exits.
// With a caller workspace the tile kernel also publishes "some source has far taps" (an epoch stamp); the far pass
// then exits at once for sub-pixel flows instead of re-reading the flow.
constexpr int TILE_W = 32, TILE_W2 = TILE_W + 2;

__device__ __forceinline__ float cover_weight(float i, float pf, float pf_p1, float pf_m1) {
  return i >= pf ? __fsub_rn(pf_p1, i) : __fsub_rn(i, pf_m1);      // > 0 iff the pixel is one of the two taps of i
}
// taps of the source pixel (w, h) the 3x3 probe of the tile kernel cannot see: in bounds, non-zero weight, and more
// than one pixel away from the source on either axis.  bit 0 nw, 1 ne, 2 sw, 3 se.  Used by BOTH kernels.
__device__ __forceinline__ int far_tap_mask(const Footprint& fp, int w, int h) {
  const int ex0 = fp.x0 - w, ey0 = fp.y0 - h;                     // tap offsets from the source: ex0, ex0+1 / ey0, ey0+1
  const bool fx0 = ex0 < -1 || ex0 > 1, fx1 = ex0 + 1 < -1 || ex0 + 1 > 1;
  const bool fy0 = ey0 < -1 || ey0 > 1, fy1 = ey0 + 1 < -1 || ey0 + 1 > 1;
  const bool vx0 = fp.wx0 > 0.f, vx1 = fp.x1ok && fp.wx1 > 0.f, vy0 = fp.wy0 > 0.f, vy1 = fp.y1ok && fp.wy1 > 0.f;
  return ((vx0 && vy0 && (fx0 || fy0)) ? 1 : 0) | ((vx1 && vy0 && (fx1 || fy0)) ? 2 : 0) |
         ((vx0 && vy1 && (fx0 || fy1)) ? 4 : 0) | ((vx1 && vy1 && (fx1 || fy1)) ? 8 : 0);
}
// 16-byte async copy with zero-fill when !valid (src-size 0)
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(valid ? 16 : 0)
               : "memory");
}
__host__ __device__ inline size_t tile_smem_bytes(int R, int LC) {
  return ((size_t)(R + 2) * TILE_W2 * (LC + 1) + (size_t)R * TILE_W * 3) * 16 + (size_t)R * TILE_W * 8;
}
// + the tap slots of the prefetching variant: [2 stages][4 taps][256 threads] x 16 B
constexpr size_t TILE_SLOT_BYTES = 2 * 4 * 256 * 16;
constexpr int ct_log2(int v) { return v <= 1 ? 0 : 1 + ct_log2(v >> 1); }

// CT / WT: channel count and row width as compile-time constants (0 = runtime): global addresses are then one base
// register plus immediates and the item index arithmetic is shifts by constants.
// PF: the 4 taps of x of the NEXT item are already on their way into per-thread shared-memory slots (cp.async, no registers
// held while in flight) and its pass-through vector into a register while the current item is computed — the loads of
// phase 2 were where 1/3 of all stall samples of the non-prefetching kernel sat (ncu source view, profiles/r2_notes.md).
template <int CT, int WT, bool PF>
__global__ void __launch_bounds__(256, PF ? 3 : 4)
warp_bwd_ndhwc_tile_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                           int64_t sB, const float* __restrict__ flow, const float* __restrict__ xs,
                           const float* __restrict__ ys, float* __restrict__ gx1, float* __restrict__ gx2,
                           float* __restrict__ gflow, int C_, int H, int W_, int lshift_, int R, int tiles_x,
                           int tiles_y, int wshift, int hshift, int* __restrict__ far_flag, int epoch) {
  extern __shared__ float4 smem4[];
  const int C = CT ? CT : C_, W = WT ? WT : W_;
  const int lshift = CT ? (ct_log2(CT / 4) < 3 ? ct_log2(CT / 4) : 3) : lshift_;     // lanes per pixel LC = min(C/4, 8)
  const int LC = 1 << lshift;
  const int nchunks = (C >> 2) >> lshift;                                                // 32-channel chunks
  const int HW = H * W;
  float4* gws = smem4;                                    // [(R+2)][34][LC] staged warped-slot gradient
  float4* sc = gws + (R + 2) * TILE_W2 * LC;              // [(R+2)][34]     source coordinates
  float4* wt = sc + (R + 2) * TILE_W2;                    // [R*32][3]       probe weights
  float2* gacc = reinterpret_cast<float2*>(wt + R * TILE_W * 3);   // [R*32] flow-gradient sums across chunks
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int bt = bid / tiles_y;
  const int h0 = ty * R, w0 = tx * TILE_W;
  const int b = bt >> 1, t = bt & 1;
  const int64_t fbase = ((int64_t)(b * 2) * 2 + t) * HW;
  const float* gw = gout + ((int64_t)(b * 4 + 1 + t) * HW) * C;
  const int n_stage = ((R + 2) * TILE_W2) << lshift;

  // staging walks (row, column) incrementally: 256 threads advance by 256/LC pixels per step, no divisions
  const int adv = 256 >> lshift, adv_r = adv / TILE_W2, adv_c = adv - adv_r * TILE_W2;
  auto stage = [&](int chunk) {
    const int sp0 = threadIdx.x >> lshift, lv = threadIdx.x & (LC - 1);
    int r = sp0 / TILE_W2, c = sp0 - r * TILE_W2;
    const float* g0 = gw + ((chunk << lshift) + lv) * 4;
    for (int i = threadIdx.x; i < n_stage; i += 256) {
      const int h = h0 - 1 + r, w = w0 - 1 + c;
      const bool ok = h >= 0 && h < H && w >= 0 && w < W;
      cp_async16_zfill(gws + i, g0 + (ok ? (h * W + w) * C : 0), ok);
      r += adv_r; c += adv_c;
      if (c >= TILE_W2) { c -= TILE_W2; ++r; }
    }
    cp_async_commit();
  };
  stage(0);

  // ---- phase 0: source coordinates ----
  bool found_far = false;
  {
    const float* fxp = flow + fbase;
    const float* fyp = fxp + 2 * (int64_t)HW;
    const float half_w = __fmul_rn((float)(W - 1), 0.5f), half_h = __fmul_rn((float)(H - 1), 0.5f);
    for (int s = threadIdx.x; s < (R + 2) * TILE_W2; s += 256) {
      const int r = s / TILE_W2;
      const int w = w0 - 1 + s - r * TILE_W2, h = h0 - 1 + r;
      float4 c = make_float4(-4.f, -4.f, 0.f, 0.f);            // outside the image: covers nothing
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(fxp + h * W + w), __ldg(fyp + h * W + w), W, H);
        c.x = __fadd_rn((float)fp.x0, fp.wx1);               // = ix exactly (wx1 = ix - floor(ix) is exact)
        c.y = __fadd_rn((float)fp.y0, fp.wy1);
        c.z = __fmul_rn(fp.gx_gate, half_w);
        c.w = __fmul_rn(fp.gy_gate, half_h);
        if (far_flag != nullptr && far_tap_mask(fp, w, h) != 0) found_far = true;
      }
      sc[s] = c;
    }
  }
  // tell the far pass (stream-ordered after this kernel) that it has work; every writer stores the same epoch
  if (found_far) *far_flag = epoch;
  __syncthreads();
  // set-up of phase 2 (and, PF, the loads of its first item) before phase 1
  const int n_items = (R * TILE_W) << lshift;
  const int rowC = W * C;
  const float* gpass = gout + ((int64_t)(b * 4 + (t ? 3 : 0)) * HW) * C;
  const float* src = (t ? x2 : x1) + b * sB;
  float* dst = (t ? gx2 : gx1) + b * sB;
  const float inv_w = 1.f / (float)W, inv_h = 1.f / (float)H;
  const int rstride = TILE_W2 * LC;                          // float4 per staged row
  const int iters = (n_items + 255) >> 8;                    // per chunk
  // prefetch state (PF only): item (pf_chunk, pf_it) goes to slot stage pf_n & 1
  float4* slots = reinterpret_cast<float4*>(gacc + R * TILE_W) + threadIdx.x;
  int pf_chunk = 0, pf_it = 0, pf_n = 0;
  float4 pass_next = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&]() {
    const int it = pf_it * 256 + threadIdx.x;
    const int pl = it >> lshift, lv = it & (LC - 1);
    const int hl = pl >> 5, wl = pl & 31;
    const int h = h0 + hl, w = w0 + wl;
    if (it < n_items && h < H && w < W) {
      const float4 own = sc[(hl + 1) * TILE_W2 + wl + 1];
      const int x0 = (int)own.x, y0 = (int)own.y;            // coordinates are clipped to [0, size-1]: truncation = floor
      const bool x1ok = x0 + 1 <= W - 1, y1ok = y0 + 1 <= H - 1;
      const int cb = (pf_chunk << lshift) * 4 + lv * 4;
      const float* xp = src + ((y0 * W + x0) * C + cb);
      float4* sl = slots + (pf_n & 1) * 4 * 256;
      cp_async16_zfill(sl, xp, true);
      cp_async16_zfill(sl + 256, x1ok ? xp + C : xp, x1ok);
      cp_async16_zfill(sl + 512, y1ok ? xp + rowC : xp, y1ok);
      cp_async16_zfill(sl + 768, (x1ok && y1ok) ? xp + rowC + C : xp, x1ok && y1ok);
      pass_next = __ldg(reinterpret_cast<const float4*>(gpass + (h * W + w) * C + cb));
    }
    cp_async_commit();
    ++pf_n;
    if (++pf_it == iters) { pf_it = 0; ++pf_chunk; }
  };
  if constexpr (PF) prefetch();                              // item 0: in flight during the rest of the prologue
  // ---- phase 1: probe weights, one thread per target pixel ----
  for (int pl = threadIdx.x; pl < R * TILE_W; pl += 256) {
    const int hl = pl >> 5, wl = pl & 31;
    const float pxf = (float)(w0 + wl), pyf = (float)(h0 + hl);
    const float pxp = pxf + 1.f, pxm = pxf - 1.f, pyp = pyf + 1.f, pym = pyf - 1.f;
    float wv[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float4* row = sc + (hl + dy) * TILE_W2 + wl;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float2 c = *reinterpret_cast<const float2*>(row + dx);
        const float wx = cover_weight(c.x, pxf, pxp, pxm), wy = cover_weight(c.y, pyf, pyp, pym);
        wv[dy * 3 + dx] = (wx > 0.f && wy > 0.f) ? __fmul_rn(wx, wy) : 0.f;
      }
    }
    wt[pl * 3 + 0] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    wt[pl * 3 + 1] = make_float4(wv[4], wv[5], wv[6], wv[7]);
    wt[pl * 3 + 2] = make_float4(wv[8], 0.f, 0.f, 0.f);
    gacc[pl] = make_float2(0.f, 0.f);
  }
  // ---- phase 2: one thread per (target pixel, channel vector), 32 channels at a time ----
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    if (chunk > 0) { __syncthreads(); stage(chunk); }
    cp_async_wait_all();                                     // the staged tile (and, PF, the slots of this chunk's first item)
    __syncthreads();
    const int cbase = (chunk << lshift) * 4;
    for (int itn = 0; itn < iters; ++itn) {
      const int it = (itn << 8) + threadIdx.x;
      const int pl = it >> lshift, lv = it & (LC - 1);
      const int hl = pl >> 5, wl = pl & 31;
      const int h = h0 + hl, w = w0 + wl;
      const bool live = it < n_items && h < H && w < W;
      float gix = 0.f, giy = 0.f;
      float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 pass = pass_next, xa, xb, xc, xd;
      if constexpr (PF) {
        const int q = chunk * iters + itn;
        const bool has_next = q + 1 < iters * nchunks;       // uniform
        if (has_next) prefetch();
        if (itn > 0) {                                       // item q's group is the last but one (or the last)
          if (has_next) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else cp_async_wait_all();
        }
        const float4* sl = slots + (q & 1) * 4 * 256;
        xa = sl[0]; xb = sl[256]; xc = sl[512]; xd = sl[768];
      }
      if (live) {
        const float4 wa = wt[pl * 3], wb = wt[pl * 3 + 1];
        const float w8 = wt[pl * 3 + 2].x;
        own = sc[(hl + 1) * TILE_W2 + wl + 1];
        const float x0f = floorf(own.x), y0f = floorf(own.y);
        const float wx0 = __fsub_rn(__fadd_rn(x0f, 1.f), own.x), wx1 = __fsub_rn(own.x, x0f);
        const float wy0 = __fsub_rn(__fadd_rn(y0f, 1.f), own.y), wy1 = __fsub_rn(own.y, y0f);
        const int eo = (h * W + w) * C + cbase + lv * 4;
        if constexpr (!PF) {
          const int x0 = (int)x0f, y0 = (int)y0f;
          const bool x1ok = x0 + 1 <= W - 1, y1ok = y0 + 1 <= H - 1;
          const float* xp = src + ((y0 * W + x0) * C + cbase + lv * 4);
          const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
          // global loads first (pass-through + the 4 taps of x), then the staged 3x3 gather
          pass = __ldg(reinterpret_cast<const float4*>(gpass + eo));
          xa = __ldg(reinterpret_cast<const float4*>(xp));
          xb = x1ok ? __ldg(reinterpret_cast<const float4*>(xp + C)) : zero4;
          xc = y1ok ? __ldg(reinterpret_cast<const float4*>(xp + rowC)) : zero4;
          xd = (x1ok && y1ok) ? __ldg(reinterpret_cast<const float4*>(xp + rowC + C)) : zero4;
        }
        const float4* gc = gws + ((hl + 1) * TILE_W2 + wl + 1) * LC + lv;        // own pixel in the staged tile
        const float wgt[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, w8};
        const float4 g = gc[0];                              // own pixel: the source-side gradient
        float4 sum = make_float4(wgt[4] * g.x, wgt[4] * g.y, wgt[4] * g.z, wgt[4] * g.w);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          if (j != 4 && wgt[j] != 0.f) {                     // a zero weight costs no shared-memory wavefront
            const float4 v = gc[(j / 3 - 1) * rstride + (j % 3 - 1) * LC];
            sum.x = fmaf(wgt[j], v.x, sum.x); sum.y = fmaf(wgt[j], v.y, sum.y);
            sum.z = fmaf(wgt[j], v.z, sum.z); sum.w = fmaf(wgt[j], v.w, sum.w);
          }
        }
        *reinterpret_cast<float4*>(dst + eo) =
            make_float4(__fadd_rn(pass.x, sum.x), __fadd_rn(pass.y, sum.y), __fadd_rn(pass.z, sum.z), __fadd_rn(pass.w, sum.w));
        // source side (out-of-bounds taps were loaded as zeros: they add exact zeros, like ATen's skipped taps)
        auto tap = [&](const float4& xv, float dgx, float dgy) {
          const float dot = fmaf(xv.w, g.w, fmaf(xv.z, g.z, fmaf(xv.y, g.y, __fmul_rn(xv.x, g.x))));
          gix = fmaf(dgx, dot, gix);
          giy = fmaf(dgy, dot, giy);
        };
        tap(xa, -wy0, -wx0);
        tap(xb, wy0, -wx1);
        tap(xc, -wy1, wx0);
        tap(xd, wy1, wx1);
      }
      for (int d = LC >> 1; d > 0; d >>= 1) {
        gix += __shfl_xor_sync(0xffffffffu, gix, d);
        giy += __shfl_xor_sync(0xffffffffu, giy, d);
      }
      if (live && lv == 0) {
        if (nchunks > 1) {                                   // channels ascending across chunks, like one long sum
          const float2 a = gacc[pl];
          gix += a.x; giy += a.y;
          gacc[pl] = make_float2(gix, giy);
        }
        if (chunk == nchunks - 1) {
          const float a = __fmul_rn(own.z, gix), c = __fmul_rn(own.w, giy);
          float* gf = gflow + fbase + h * W + w;
          gf[0] = wshift >= 0 ? __fmul_rn(a, inv_w) : __fdiv_rn(a, (float)W);       // exact for power-of-two sizes
          gf[2 * (int64_t)HW] = hshift >= 0 ? __fmul_rn(c, inv_h) : __fdiv_rn(c, (float)H);
        }
      }
    }
  }
}

// far taps of the tile gather: one lane per source pixel-frame finds them, the warp then adds them cooperatively
__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_far_kernel(const float* __restrict__ gout, int64_t sB, const float* __restrict__ flow,
                          const float* __restrict__ xs, const float* __restrict__ ys, float* __restrict__ gx1,
                          float* __restrict__ gx2, int C, int H, int W, int q, int planes,
                          const int* __restrict__ far_flag, int epoch) {
  if (far_flag != nullptr && *far_flag != epoch) return;       // the tile kernel saw no far tap in this launch
  const int HW = H * W;
  const int64_t total = (int64_t)HW * planes, stride = (int64_t)gridDim.x * 256;
  const int64_t rounds = (total + stride - 1) / stride;
  const int lane = threadIdx.x & 31;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (int64_t r = 0; r < rounds; ++r, i += stride) {
    int mask = 0, bt = 0, p = 0, o_nw = 0;
    float wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f;
    if (i < total) {
      bt = (int)(i / HW);
      p = (int)(i - (int64_t)bt * HW);
      const int h = p / W, w = p - h * W;
      const int64_t fo = ((int64_t)(bt >> 1) * 4 + (bt & 1)) * HW + p;
      const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
      wx0 = fp.wx0; wx1 = fp.wx1; wy0 = fp.wy0; wy1 = fp.wy1;
      o_nw = fp.y0 * W + fp.x0;
      mask = far_tap_mask(fp, w, h);
    }
    const unsigned any = __ballot_sync(0xffffffffu, mask != 0);
    if (any == 0) continue;
    // the warp serves `nsub` sources at a time: with q = C/4 < 32 channel vectors per pixel, 32/q sources share the lanes
    const int nsub = q >= 32 ? 1 : 32 / q, qq = q >= 32 ? 32 : q;
    const int sub = lane / qq, v0 = lane - sub * qq;
    for (int base = 0; base < 32; base += nsub) {
      if (((any >> base) & ((nsub == 32 ? 0xffffffffu : ((1u << nsub) - 1u)))) == 0) continue;
      const int sl = base + sub;
      const int m = __shfl_sync(0xffffffffu, mask, sl), sbt = __shfl_sync(0xffffffffu, bt, sl);
      const int sp = __shfl_sync(0xffffffffu, p, sl), so = __shfl_sync(0xffffffffu, o_nw, sl);
      const float a0 = __shfl_sync(0xffffffffu, wx0, sl), a1 = __shfl_sync(0xffffffffu, wx1, sl);
      const float b0 = __shfl_sync(0xffffffffu, wy0, sl), b1 = __shfl_sync(0xffffffffu, wy1, sl);
      if (m == 0) continue;
      const int b = sbt >> 1, t = sbt & 1;
      const float* gwp = gout + ((int64_t)(b * 4 + 1 + t) * HW + sp) * C;
      float* dst = (t ? gx2 : gx1) + b * sB;
      for (int v = v0; v < q; v += 32) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gwp + v * 4));
        auto add = [&](int off, float wgt) {
          atomicAdd(reinterpret_cast<float4*>(dst + (int64_t)off * C + v * 4), make_float4(wgt * g.x, wgt * g.y, wgt * g.z, wgt * g.w));
        };
        if (m & 1) add(so, __fmul_rn(a0, b0));
        if (m & 2) add(so + 1, __fmul_rn(a1, b0));
        if (m & 4) add(so + W, __fmul_rn(a0, b1));
        if (m & 8) add(so + W + 1, __fmul_rn(a1, b1));
      }
    }
  }
}

// ---- bf16 storage: the same single-pass tile gather with 8-channel vectors -------------------------------------------
// Features and gradients are bf16 in HBM (half the bytes), every sum runs in fp32: the staged warped-slot gradient
// tile holds raw bf16 (16 bytes = 8 channels per unit), phase 2 widens to fp32 registers, accumulates the pass-through
// slot + the 3x3 probe in fp32 and rounds ONCE on the way out.  Coordinates, probe weights and the flow gradient are
// the fp32 code of the kernel above.  Far taps are added by bf16x2 reductions (each rounds to bf16: fine at 2e-2).
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}

__global__ void __launch_bounds__(256, 4)
warp_bwd_ndhwc_tile_bf16_kernel(const __nv_bfloat16* __restrict__ gout, const __nv_bfloat16* __restrict__ x1,
                                const __nv_bfloat16* __restrict__ x2, int64_t sB, const float* __restrict__ flow,
                                const float* __restrict__ xs, const float* __restrict__ ys, __nv_bfloat16* __restrict__ gx1,
                                __nv_bfloat16* __restrict__ gx2, float* __restrict__ gflow, int C, int H, int W, int lshift,
                                int R, int tiles_x, int tiles_y, int wshift, int hshift, int* __restrict__ far_flag, int epoch) {
  extern __shared__ float4 smem4[];
  const int LC = 1 << lshift;                               // lanes per pixel = min(C/8, 8)
  const int nchunks = (C >> 3) >> lshift;                    // chunks of LC*8 channels
  const int HW = H * W;
  uint4* gws = reinterpret_cast<uint4*>(smem4);             // [(R+2)][34][LC] staged warped-slot gradient (raw bf16)
  float4* sc = smem4 + (R + 2) * TILE_W2 * LC;              // [(R+2)][34]     source coordinates
  float4* wt = sc + (R + 2) * TILE_W2;                      // [R*32][3]       probe weights
  float2* gacc = reinterpret_cast<float2*>(wt + R * TILE_W * 3);
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int bt = bid / tiles_y;
  const int h0 = ty * R, w0 = tx * TILE_W;
  const int b = bt >> 1, t = bt & 1;
  const int64_t fbase = ((int64_t)(b * 2) * 2 + t) * HW;
  const __nv_bfloat16* gw = gout + ((int64_t)(b * 4 + 1 + t) * HW) * C;
  const int n_stage = ((R + 2) * TILE_W2) << lshift;
  const int adv = 256 >> lshift, adv_r = adv / TILE_W2, adv_c = adv - adv_r * TILE_W2;
  auto stage = [&](int chunk) {
    const int sp0 = threadIdx.x >> lshift, lv = threadIdx.x & (LC - 1);
    int r = sp0 / TILE_W2, c = sp0 - r * TILE_W2;
    const __nv_bfloat16* g0 = gw + ((chunk << lshift) + lv) * 8;
    for (int i = threadIdx.x; i < n_stage; i += 256) {
      const int h = h0 - 1 + r, w = w0 - 1 + c;
      const bool ok = h >= 0 && h < H && w >= 0 && w < W;
      cp_async16_zfill(gws + i, g0 + (ok ? (h * W + w) * C : 0), ok);
      r += adv_r; c += adv_c;
      if (c >= TILE_W2) { c -= TILE_W2; ++r; }
    }
    cp_async_commit();
  };
  stage(0);
  bool found_far = false;
  {
    const float* fxp = flow + fbase;
    const float* fyp = fxp + 2 * (int64_t)HW;
    const float half_w = __fmul_rn((float)(W - 1), 0.5f), half_h = __fmul_rn((float)(H - 1), 0.5f);
    for (int s = threadIdx.x; s < (R + 2) * TILE_W2; s += 256) {
      const int r = s / TILE_W2;
      const int w = w0 - 1 + s - r * TILE_W2, h = h0 - 1 + r;
      float4 c = make_float4(-4.f, -4.f, 0.f, 0.f);
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(fxp + h * W + w), __ldg(fyp + h * W + w), W, H);
        c.x = __fadd_rn((float)fp.x0, fp.wx1);
        c.y = __fadd_rn((float)fp.y0, fp.wy1);
        c.z = __fmul_rn(fp.gx_gate, half_w);
        c.w = __fmul_rn(fp.gy_gate, half_h);
        if (far_flag != nullptr && far_tap_mask(fp, w, h) != 0) found_far = true;
      }
      sc[s] = c;
    }
  }
  if (found_far) *far_flag = epoch;
  __syncthreads();
  for (int pl = threadIdx.x; pl < R * TILE_W; pl += 256) {
    const int hl = pl >> 5, wl = pl & 31;
    const float pxf = (float)(w0 + wl), pyf = (float)(h0 + hl);
    const float pxp = pxf + 1.f, pxm = pxf - 1.f, pyp = pyf + 1.f, pym = pyf - 1.f;
    float wv[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float4* row = sc + (hl + dy) * TILE_W2 + wl;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float2 c = *reinterpret_cast<const float2*>(row + dx);
        const float wx = cover_weight(c.x, pxf, pxp, pxm), wy = cover_weight(c.y, pyf, pyp, pym);
        wv[dy * 3 + dx] = (wx > 0.f && wy > 0.f) ? __fmul_rn(wx, wy) : 0.f;
      }
    }
    wt[pl * 3 + 0] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    wt[pl * 3 + 1] = make_float4(wv[4], wv[5], wv[6], wv[7]);
    wt[pl * 3 + 2] = make_float4(wv[8], 0.f, 0.f, 0.f);
    gacc[pl] = make_float2(0.f, 0.f);
  }
  const int n_items = (R * TILE_W) << lshift;
  const int rowC = W * C;
  const __nv_bfloat16* gpass = gout + ((int64_t)(b * 4 + (t ? 3 : 0)) * HW) * C;
  const __nv_bfloat16* src = (t ? x2 : x1) + b * sB;
  __nv_bfloat16* dst = (t ? gx2 : gx1) + b * sB;
  const float inv_w = 1.f / (float)W, inv_h = 1.f / (float)H;
  const int rstride = TILE_W2 * LC;
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    if (chunk > 0) { __syncthreads(); stage(chunk); }
    cp_async_wait_all();
    __syncthreads();
    const int cbase = (chunk << lshift) * 8;
    for (int base = 0; base < n_items; base += 256) {
      const int it = base + threadIdx.x;
      const int pl = it >> lshift, lv = it & (LC - 1);
      const int hl = pl >> 5, wl = pl & 31;
      const int h = h0 + hl, w = w0 + wl;
      const bool live = it < n_items && h < H && w < W;
      float gix = 0.f, giy = 0.f;
      float4 own = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const float4 wa = wt[pl * 3], wb = wt[pl * 3 + 1];
        const float w8 = wt[pl * 3 + 2].x;
        own = sc[(hl + 1) * TILE_W2 + wl + 1];
        const float x0f = floorf(own.x), y0f = floorf(own.y);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const float wx0 = __fsub_rn(__fadd_rn(x0f, 1.f), own.x), wx1 = __fsub_rn(own.x, x0f);
        const float wy0 = __fsub_rn(__fadd_rn(y0f, 1.f), own.y), wy1 = __fsub_rn(own.y, y0f);
        const bool x1ok = x0 + 1 <= W - 1, y1ok = y0 + 1 <= H - 1;
        const int eo = (h * W + w) * C + cbase + lv * 8;
        const __nv_bfloat16* xp = src + ((y0 * W + x0) * C + cbase + lv * 8);
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        const uint4 pass = __ldg(reinterpret_cast<const uint4*>(gpass + eo));
        const uint4 xa = __ldg(reinterpret_cast<const uint4*>(xp));
        const uint4 xb = x1ok ? __ldg(reinterpret_cast<const uint4*>(xp + C)) : z4;
        const uint4 xc = y1ok ? __ldg(reinterpret_cast<const uint4*>(xp + rowC)) : z4;
        const uint4 xd = (x1ok && y1ok) ? __ldg(reinterpret_cast<const uint4*>(xp + rowC + C)) : z4;
        const uint4* gc = gws + ((hl + 1) * TILE_W2 + wl + 1) * LC + lv;
        const float wgt[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, w8};
        float g[8], sum[8], tmp[8];
        unpack8(gc[0], g);
#pragma unroll
        for (int k = 0; k < 8; ++k) sum[k] = wgt[4] * g[k];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          if (j != 4 && wgt[j] != 0.f) {
            unpack8(gc[(j / 3 - 1) * rstride + (j % 3 - 1) * LC], tmp);
#pragma unroll
            for (int k = 0; k < 8; ++k) sum[k] = fmaf(wgt[j], tmp[k], sum[k]);
          }
        }
        unpack8(pass, tmp);
        Pack<__nv_bfloat16> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.f[k] = __fadd_rn(tmp[k], sum[k]);
        st_pack<__nv_bfloat16>(dst + eo, o);
        auto tap = [&](const uint4& xv, float dgx, float dgy) {
          unpack8(xv, tmp);
          float dot = __fmul_rn(tmp[0], g[0]);
#pragma unroll
          for (int k = 1; k < 8; ++k) dot = fmaf(tmp[k], g[k], dot);
          gix = fmaf(dgx, dot, gix);
          giy = fmaf(dgy, dot, giy);
        };
        tap(xa, -wy0, -wx0);
        tap(xb, wy0, -wx1);
        tap(xc, -wy1, wx0);
        tap(xd, wy1, wx1);
      }
      for (int d = LC >> 1; d > 0; d >>= 1) {
        gix += __shfl_xor_sync(0xffffffffu, gix, d);
        giy += __shfl_xor_sync(0xffffffffu, giy, d);
      }
      if (live && lv == 0) {
        if (nchunks > 1) {
          const float2 a = gacc[pl];
          gix += a.x; giy += a.y;
          gacc[pl] = make_float2(gix, giy);
        }
        if (chunk == nchunks - 1) {
          const float a = __fmul_rn(own.z, gix), c = __fmul_rn(own.w, giy);
          float* gf = gflow + fbase + h * W + w;
          gf[0] = wshift >= 0 ? __fmul_rn(a, inv_w) : __fdiv_rn(a, (float)W);
          gf[2 * (int64_t)HW] = hshift >= 0 ? __fmul_rn(c, inv_h) : __fdiv_rn(c, (float)H);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_far_bf16_kernel(const __nv_bfloat16* __restrict__ gout, int64_t sB, const float* __restrict__ flow,
                               const float* __restrict__ xs, const float* __restrict__ ys, __nv_bfloat16* __restrict__ gx1,
                               __nv_bfloat16* __restrict__ gx2, int C, int H, int W, int q, int planes,
                               const int* __restrict__ far_flag, int epoch) {
  if (far_flag != nullptr && *far_flag != epoch) return;
  const int HW = H * W;
  const int64_t total = (int64_t)HW * planes, stride = (int64_t)gridDim.x * 256;
  const int64_t rounds = (total + stride - 1) / stride;
  const int lane = threadIdx.x & 31;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (int64_t r = 0; r < rounds; ++r, i += stride) {
    int mask = 0, bt = 0, p = 0, o_nw = 0;
    float wx0 = 0.f, wx1 = 0.f, wy0 = 0.f, wy1 = 0.f;
    if (i < total) {
      bt = (int)(i / HW);
      p = (int)(i - (int64_t)bt * HW);
      const int h = p / W, w = p - h * W;
      const int64_t fo = ((int64_t)(bt >> 1) * 4 + (bt & 1)) * HW + p;
      const Footprint fp = footprint_auto(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
      wx0 = fp.wx0; wx1 = fp.wx1; wy0 = fp.wy0; wy1 = fp.wy1;
      o_nw = fp.y0 * W + fp.x0;
      mask = far_tap_mask(fp, w, h);
    }
    const unsigned any = __ballot_sync(0xffffffffu, mask != 0);
    if (any == 0) continue;
    const int nsub = q >= 32 ? 1 : 32 / q, qq = q >= 32 ? 32 : q;
    const int sub = lane / qq, v0 = lane - sub * qq;
    for (int base = 0; base < 32; base += nsub) {
      if (((any >> base) & ((nsub == 32 ? 0xffffffffu : ((1u << nsub) - 1u)))) == 0) continue;
      const int sl = base + sub;
      const int m = __shfl_sync(0xffffffffu, mask, sl), sbt = __shfl_sync(0xffffffffu, bt, sl);
      const int sp = __shfl_sync(0xffffffffu, p, sl), so = __shfl_sync(0xffffffffu, o_nw, sl);
      const float a0 = __shfl_sync(0xffffffffu, wx0, sl), a1 = __shfl_sync(0xffffffffu, wx1, sl);
      const float b0 = __shfl_sync(0xffffffffu, wy0, sl), b1 = __shfl_sync(0xffffffffu, wy1, sl);
      if (m == 0) continue;
      const int b = sbt >> 1, t = sbt & 1;
      const __nv_bfloat16* gwp = gout + ((int64_t)(b * 4 + 1 + t) * HW + sp) * C;
      __nv_bfloat16* dst = (t ? gx2 : gx1) + b * sB;
      for (int v = v0; v < q; v += 32) {
        float g[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(gwp + v * 8)), g);
        auto add = [&](int off, float wgt) {
          __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(dst + (int64_t)off * C + v * 8);
#pragma unroll
          for (int k = 0; k < 4; ++k) atomicAdd(d2 + k, __floats2bfloat162_rn(wgt * g[2 * k], wgt * g[2 * k + 1]));
        };
        if (m & 1) add(so, __fmul_rn(a0, b0));
        if (m & 2) add(so + 1, __fmul_rn(a1, b0));
        if (m & 4) add(so + W, __fmul_rn(a0, b1));
        if (m & 8) add(so + W + 1, __fmul_rn(a1, b1));
      }
    }
  }
}

// ---- (S) vector-atomic scatter; both kernels are grid-stride over (plane, item) so that the "other family did
// the work" exit costs a few hundred blocks, not one block per 256 items ----
template <typename T>
__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_init_kernel(const T* __restrict__ gout, T* __restrict__ gx1, T* __restrict__ gx2, int64_t sB,
                           float* __restrict__ gflow, int C, int HW, int q, int planes, bool zero_gflow,
                           const int* __restrict__ hdr) {
  constexpr int V = CVec<T>::N;
  if (hdr != nullptr && gather_active(hdr)) return;
  const int64_t per_plane = (int64_t)HW * q, total = per_plane * planes;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int bt = (int)(i / per_plane);
    const int idx = (int)(i - (int64_t)bt * per_plane);
    const int b = bt >> 1, t = bt & 1;
    const int64_t e = (int64_t)idx * V;                        // element offset inside one (b, frame) plane
    *reinterpret_cast<uint4*>((t ? gx2 : gx1) + b * sB + e) =
        __ldg(reinterpret_cast<const uint4*>(gout + ((int64_t)(b * 4 + (t ? 3 : 0)) * HW) * C + e));
    if (zero_gflow && idx < HW) {
      const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + idx;
      gflow[fo] = 0.f; gflow[fo + 2 * (int64_t)HW] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(256)
warp_bwd_ndhwc_scatter_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                              int64_t sB, const float* __restrict__ flow, const float* __restrict__ xs,
                              const float* __restrict__ ys, float* __restrict__ gx1, float* __restrict__ gx2,
                              float* __restrict__ gflow, int C, int H, int W, int q, int qshift, int planes,
                              bool shuffle_reduce, const int* __restrict__ hdr) {
  if (hdr != nullptr && gather_active(hdr)) return;
  const int HW = H * W;
  const int64_t per_plane = (int64_t)HW * q, total = per_plane * planes;
  // every lane of a warp runs the same number of iterations (total is padded to the stride by the `live` flag)
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int64_t rounds = (total + stride - 1) / stride;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (int64_t r = 0; r < rounds; ++r, i += stride) {
    const bool live = i < total;
    int p = 0, v = 0, b = 0, t = 0;
    if (live) {
      const int bt = (int)(i / per_plane);
      split_item((int)(i - (int64_t)bt * per_plane), q, qshift, p, v);
      b = bt >> 1; t = bt & 1;
    }
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
}

static int ilog2_exact(int q) {
  for (int s = 0; s < 31; ++s)
    if ((1 << s) == q) return s;
  return -1;
}

int64_t warp_bwd_ndhwc_workspace_bytes(int B, int H, int W) { return bwd_ws_bytes((int64_t)B * 2 * H * W); }

template <typename T>
int warp_fwd_ndhwc(const T* x1, const T* x2, int64_t sB, const float* flow, const float* xs, const float* ys, T* out,
                   int B, int C, int H, int W, cudaStream_t st) {
  constexpr int V = CVec<T>::N;
  if (C % V || !aligned16(x1) || !aligned16(x2) || !aligned16(out) || sB % V)
    return fail(SMOW_EALIGN, "NDHWC warp needs C %% %d == 0 and 16 B aligned tensors", V);
  const int q = C / V;
  if ((int64_t)H * W * q >= (1ll << 31)) return fail(SMOW_ERANGE, "plane too large");
  const int qs = ilog2_exact(q);
  if (qs >= 0 && q <= 2048 && option(OPT_WARP_FWD_VARIANT) != 0) {        // default: per-pixel coordinates, shuffled
    const int qg = q <= 32 ? q : 32;                                        // vectors per pixel handled by one warp pass
    dim3 grid((unsigned)(((int64_t)H * W + 255) / 256), 2 * B, q / qg);
    if constexpr (std::is_same<T, float>::value)
      warp_fwd_ndhwc_shfl_kernel<T><<<grid, 256, 0, st>>>(x1, x2, sB, flow, xs, ys, out, C, H, W, qg, ilog2_exact(qg), ilog2_exact(W));
    else
      warp_fwd_ndhwc_shfl4_kernel<T><<<grid, 256, 0, st>>>(x1, x2, sB, flow, xs, ys, out, C, H, W, qg, ilog2_exact(qg), ilog2_exact(W));
  } else {                                                                 // any q; also warp_fwd_variant = 0
    dim3 grid((unsigned)(((int64_t)H * W * q + 255) / 256), 2 * B);
    warp_fwd_ndhwc_kernel<T><<<grid, 256, 0, st>>>(x1, x2, sB, flow, xs, ys, out, C, H, W, q, qs);
  }
  count_launch();
  return check_launch("warp_fwd_ndhwc");
}

template <typename T>
int warp_bwd_ndhwc(const T* gout, const T* x1, const T* x2, int64_t sB, const float* flow, const float* xs,
                   const float* ys, T* gx1, T* gx2, float* gflow, int B, int C, int H, int W, void* ws,
                   int64_t ws_bytes, cudaStream_t st) {
  if constexpr (!std::is_same<T, float>::value) {
    // bf16 storage: single-pass tile gather with 8-channel vectors + bf16x2 far-tap reductions
    if (C % 8 || !aligned16(gout) || !aligned16(x1) || !aligned16(x2) || !aligned16(gx1) || !aligned16(gx2) || sB % 8)
      return fail(SMOW_EALIGN, "NDHWC bf16 warp needs C %% 8 == 0 and 16 B aligned tensors");
    const int q = C / 8, qs = ilog2_exact(q);
    const int HW = H * W;
    if (qs < 0) return fail(SMOW_EDTYPE, "NDHWC bf16 warp backward needs C / 8 to be a power of two (got C = %d)", C);
    if ((int64_t)HW * C >= (1ll << 31)) return fail(SMOW_ERANGE, "plane too large");
    const int lshift = qs < 3 ? qs : 3;
    int R = option(OPT_NDHWC_BWD_ROWS);
    // measured on B200 (benchmarks/bwd_bf16_probe.py) with 4 CTAs per SM (64 registers, r2: 104 registers / 2 CTAs gave 0.30-0.34
    // of the roofline, now 0.43-0.45): 6-row tiles for >= 64 channels, 8 rows at 32, 12 at 16
    if (R <= 0) R = lshift == 3 ? 6 : lshift == 2 ? 8 : 12;
    if (R > H) R = H;
    const size_t smem = tile_smem_bytes(R, 1 << lshift);
    if (smem > 72 * 1024) return fail(SMOW_ERANGE, "NDHWC bf16 warp backward: tile does not fit shared memory");
    const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + R - 1) / R;
    const unsigned tgrid = (unsigned)(tiles_x * tiles_y * 2 * B);
    static std::atomic<int> g_epoch16{1 << 20};
    int* far_flag = (ws != nullptr && ws_bytes >= 64 && aligned16(ws)) ? reinterpret_cast<int*>(ws) : nullptr;
    const int epoch = g_epoch16.fetch_add(1, std::memory_order_relaxed) + 1;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(warp_bwd_ndhwc_tile_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    warp_bwd_ndhwc_tile_bf16_kernel<<<tgrid, 256, smem, st>>>(gout, x1, x2, sB, flow, xs, ys, gx1, gx2, gflow, C, H, W, lshift, R,
                                                              tiles_x, tiles_y, ilog2_exact(W), ilog2_exact(H), far_flag, epoch);
    const int64_t pf = (int64_t)HW * 2 * B;
    const int fcap = device_info().sms * 4;
    const int fgrid = (int)((pf + 255) / 256 < fcap ? (pf + 255) / 256 : fcap);
    warp_bwd_ndhwc_far_bf16_kernel<<<fgrid, 256, 0, st>>>(gout, sB, flow, xs, ys, gx1, gx2, C, H, W, q, 2 * B, far_flag, epoch);
    count_launch(2);
    return check_launch("warp_bwd_ndhwc_tile (bf16)");
  } else {
    if (C % 4 || !aligned16(gout) || !aligned16(x1) || !aligned16(x2) || !aligned16(gx1) || !aligned16(gx2) || sB % 4)
      return fail(SMOW_EALIGN, "NDHWC warp needs C %% 4 == 0 and 16 B aligned tensors");
    const int q = C / 4, qs = ilog2_exact(q);
    const int HW = H * W;
    if ((int64_t)HW * q >= (1ll << 31)) return fail(SMOW_ERANGE, "plane too large");
    const bool shuffle = qs >= 0;
    const int64_t n_pf = (int64_t)B * 2 * HW;
    dim3 grid((unsigned)(((int64_t)HW * q + 255) / 256), 2 * B);
    int* hdr = nullptr;
    int launches = 0;
    // (G) only on request (warp_bwd_variant = 3: bit-reproducible gradients) and with a workspace: on B200 the
    // vector-atomic scatter is the faster of the two (profiles/r1_notes.md)
    const bool gather = option(OPT_WARP_BWD_VARIANT) == 3 && ws != nullptr && ws_bytes >= bwd_ws_bytes(n_pf) &&
                        shuffle && q <= 32 && aligned16(ws);
    if (gather) {
      const BwdWs w = carve_ws(ws, n_pf);
      hdr = w.hdr;
      dim3 pgrid((HW + 255) / 256, 2 * B);
      warp_bwd_ndhwc_hdr_kernel<<<1, 32, 0, st>>>(hdr);
      warp_bwd_ndhwc_stat_kernel<<<pgrid, 256, 0, st>>>(flow, xs, ys, w, H, W);
      warp_bwd_ndhwc_list_kernel<<<pgrid, 256, 0, st>>>(w, H, W, n_pf);
      warp_bwd_ndhwc_apply_kernel<<<grid, 256, 0, st>>>(gout, x1, x2, sB, flow, xs, ys, w, gx1, gx2, gflow, C, H, W, q, qs,
                                                        n_pf);
      launches += 4;
    }
    // (T) default: single-pass tile gather + far-tap pass (variant -1 / 4)
    const int bv = option(OPT_WARP_BWD_VARIANT);
    if (!gather && (bv < 0 || bv == 4) && shuffle) {
      int R = option(OPT_NDHWC_BWD_ROWS);
      const int lshift = qs < 3 ? qs : 3;
      // measured on B200 (benchmarks/bwd_probe.py, 4 CTAs per SM): 5-row tiles for >= 32 channels (0.71-0.73 of the roofline;
      // 4 rows: 0.70-0.72, 6 rows: 0.63-0.68), 8-row tiles below
      if (R <= 0) R = lshift == 3 ? 5 : 8;
      if (R > H) R = H;
      // next item's x taps prefetched through shared-memory slots (knob ndhwc_bwd_pf: -1 auto, 0 off, 1 on).  Measured on
      // B200 (benchmarks/bwd_probe.py): +4 % / +11 % at C = 128 / 256 (many 32-channel chunks per tile: the prefetch runs
      // across the chunk boundaries), -7 % at C <= 64 (3 instead of 4 CTAs per SM) => auto = C >= 128
      const int pfo = option(OPT_NDHWC_BWD_PF);
      const bool pfv = pfo < 0 ? C >= 128 : pfo != 0;
      const size_t smem = tile_smem_bytes(R, 1 << lshift) + (pfv ? TILE_SLOT_BYTES : 0);
      if (smem <= 104 * 1024 && (int64_t)HW * C < (1ll << 31)) {
        const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + R - 1) / R;
        const unsigned tgrid = (unsigned)(tiles_x * tiles_y * 2 * B);
        // optional caller workspace (>= 64 B, 16 B aligned): word 0 carries the far-tap epoch stamp of this call
        static std::atomic<int> g_epoch{0};
        int* far_flag = (ws != nullptr && ws_bytes >= 64 && aligned16(ws)) ? reinterpret_cast<int*>(ws) : nullptr;
        const int epoch = g_epoch.fetch_add(1, std::memory_order_relaxed) + 1;
#define SMOW_TILE_LAUNCH2(CT, WT, PFV)                                                                                \
  do {                                                                                                               \
    if (smem > 48 * 1024)                                                                                            \
      cudaFuncSetAttribute(warp_bwd_ndhwc_tile_kernel<CT, WT, PFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024); \
    warp_bwd_ndhwc_tile_kernel<CT, WT, PFV><<<tgrid, 256, smem, st>>>(gout, x1, x2, sB, flow, xs, ys, gx1, gx2, gflow, C, \
                                                                 H, W, lshift, R, tiles_x, tiles_y, ilog2_exact(W), \
                                                                 ilog2_exact(H), far_flag, epoch);                   \
  } while (0)
#define SMOW_TILE_LAUNCH(CT, WT) do { if (pfv) SMOW_TILE_LAUNCH2(CT, WT, true); else SMOW_TILE_LAUNCH2(CT, WT, false); } while (0)
        // compile-time (C, W) for the models' shapes (C = 16 / 32 at 128 x 128) and the sweep's; anything else is generic
        bool done = false;
#define SMOW_TILE_CASE(CT, WT) if (!done && C == CT && W == WT) { SMOW_TILE_LAUNCH(CT, WT); done = true; }
        SMOW_TILE_CASE(16, 128) SMOW_TILE_CASE(32, 128) SMOW_TILE_CASE(64, 128) SMOW_TILE_CASE(128, 128)
        SMOW_TILE_CASE(256, 128) SMOW_TILE_CASE(64, 64) SMOW_TILE_CASE(128, 64) SMOW_TILE_CASE(256, 64)
        SMOW_TILE_CASE(64, 256) SMOW_TILE_CASE(128, 256) SMOW_TILE_CASE(256, 256)
        if (!done) SMOW_TILE_LAUNCH(0, 0);
#undef SMOW_TILE_CASE
#undef SMOW_TILE_LAUNCH
#undef SMOW_TILE_LAUNCH2
        const int64_t pf = (int64_t)HW * 2 * B;
        const int fcap = device_info().sms * 4;       // grid-stride: a small grid exits fastest when there is no far tap
        const int fgrid = (int)((pf + 255) / 256 < fcap ? (pf + 255) / 256 : fcap);
        warp_bwd_ndhwc_far_kernel<<<fgrid, 256, 0, st>>>(gout, sB, flow, xs, ys, gx1, gx2, C, H, W, q, 2 * B, far_flag, epoch);
        count_launch(2);
        return check_launch("warp_bwd_ndhwc_tile");
      }
    }
    // (S) in batch chunks sized for the L2: a chunk's gx lines are still L2-resident when its vector reductions
    // arrive, so the read-modify-write never reaches HBM (one chunk for the models' per-GPU batches)
    // (knob bwd_chunk_mb, 0 = one chunk: measured on B200 the launch granularity costs more than the RMW saves)
    const int64_t per_pair = (int64_t)10 * C * HW * (int64_t)sizeof(float);     // pass + gx + warp-slot + x, both frames
    const int chunk_mb = option(OPT_BWD_CHUNK_MB);
    int bc = chunk_mb > 0 ? (int)((int64_t)chunk_mb * 1024 * 1024 / per_pair) : B;
    if (bc < 1) bc = 1;
    const int cap = device_info().sms * 64;
    for (int b0 = 0; b0 < B; b0 += bc) {
      const int nb = (B - b0) < bc ? (B - b0) : bc;
      const int64_t items = (int64_t)HW * q * 2 * nb;
      const int sgrid = (int)((items + 255) / 256 < cap ? (items + 255) / 256 : cap);
      const float* g0 = gout + (int64_t)b0 * 4 * HW * C;
      const float* f0 = flow + (int64_t)b0 * 4 * HW;
      float* gf0 = gflow + (int64_t)b0 * 4 * HW;
      warp_bwd_ndhwc_init_kernel<float><<<sgrid, 256, 0, st>>>(g0, gx1 + b0 * sB, gx2 + b0 * sB, sB, gf0, C, HW, q, 2 * nb,
                                                               !shuffle || q > 32, hdr);
      warp_bwd_ndhwc_scatter_kernel<<<sgrid, 256, 0, st>>>(g0, x1 + b0 * sB, x2 + b0 * sB, sB, f0, xs, ys, gx1 + b0 * sB,
                                                           gx2 + b0 * sB, gf0, C, H, W, q, qs, 2 * nb, shuffle, hdr);
      launches += 2;
    }
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
