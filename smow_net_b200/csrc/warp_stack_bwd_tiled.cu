// A1 backward, tiled inverse-gather variant (NCDHW, sm_100a).
//
//   gx[:,:,t]  = gout[:,:,pass(t)] + W_t^T gout[:,:,1+t]      (W_t = the sparse bilinear sampling matrix)
//   gflow      = d out / d flow, gated by the clamp / border-clip masks
//
// ATen evaluates W^T g as 4*C float atomics per pixel into a zero-filled buffer.  Here the
// scatter is turned into a GATHER that needs neither atomics nor a zero-fill:
//
//  * a CTA owns the gx rows [h0, h0+R) of one (pair, frame) plane and writes them exactly
//    once with plain coalesced stores;
//  * phase 0: the clipped sample coordinates (ix, iy) of every pixel in the source window
//    [h0-HALO, h0+R+HALO) go to shared memory, with the tile-wide range of integer displacements;
//  * phase 1: every thread PROBES, for each of its target pixels, the few source pixels whose
//    bilinear footprint can cover it (3x3 candidates for sub-pixel flows) and keeps the hits as a
//    (weight, source offset) list in registers — computed once per tile, reused for all channels;
//    probe order is fixed, so the result is deterministic (ATen's is not);
//  * phase 2: channel chunks stream through shared memory (1-D bulk copies + mbarrier, two
//    stages): x window, gout[warp slot] window, gout[pass slot] rows.  Per channel a thread does
//    its list's multiply-adds for gx and the 4-tap derivative sums for gflow;
//  * contributions whose source lies outside the window (|row offset| > HALO) or whose column
//    displacement exceeds DCAP are left to a second tiny launch (warp_bwd_far_kernel) that
//    re-derives the same predicate from the flow alone and adds them with L2 atomics.  For
//    sub-pixel flows it reads the flow (1/C of the traffic) and exits.
#include "warp_stack_tiled.cuh"

namespace smow {

constexpr int BWD_THREADS = 512;
constexpr int BWD_NP = 2;      // target pixels per thread
constexpr int BWD_K = 8;       // register list length per target
constexpr int BWD_CC = 4;      // channels per pipeline stage
constexpr int BWD_NSTAGE = 2;

struct BwdGeom { int R, HALO, DCAP; };

// Is the contribution of source row sy to target row ty handled by the tile kernel?
__device__ __forceinline__ bool near_row(int sy, int ty, const BwdGeom& g) {
  const int h0 = (ty / g.R) * g.R;
  return sy >= h0 - g.HALO && sy < h0 + g.R + g.HALO;
}
__device__ __forceinline__ bool near_col(int dx, const BwdGeom& g) { return dx >= -g.DCAP && dx <= g.DCAP; }

template <typename T>
__global__ void __launch_bounds__(BWD_THREADS, 1)
warp_bwd_tiled_kernel(const T* __restrict__ gout, const T* __restrict__ x1, const T* __restrict__ x2,
                      int64_t sB, int64_t sC, const float* __restrict__ flow, const float* __restrict__ xs,
                      const float* __restrict__ ys, T* __restrict__ gx1, T* __restrict__ gx2,
                      float* __restrict__ gflow, int C, int H, int W, BwdGeom geo, int nbands, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int R = geo.R, HALO = geo.HALO;
  const int WR = R + 2 * HALO;
  const int HW = H * W;
  const int plane = WR * W;     // elements per channel of a window buffer
  const int core = R * W;       // elements per channel of the pass-through buffer
  // stage layout: [x window CC*plane][gout warp-slot window CC*plane][gout pass-slot rows CC*core]
  const uint32_t stage_elems = (uint32_t)(BWD_CC * (2 * plane + core));
  const uint32_t stage_bytes = stage_elems * (uint32_t)sizeof(T);
  float* s_ix = reinterpret_cast<float*>(smem_raw + (size_t)BWD_NSTAGE * stage_bytes);
  float* s_iy = s_ix + plane;
  int* s_rng = reinterpret_cast<int*>(s_iy + plane);          // dxmin, dxmax, dymin, dymax
  uint64_t* full = reinterpret_cast<uint64_t*>(s_rng + 4);
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < BWD_NSTAGE; ++s) mbar_init(full + s, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int nchunk = C / BWD_CC;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_it = my_tiles * nchunk;

  auto issue = [&](int it) {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    const int tile = blockIdx.x + tl * gridDim.x;
    const int band = tile % nbands, bt = tile / nbands;
    const int b = bt >> 1, t = bt & 1;
    const int h0 = band * R;
    const int r_lo = max(0, h0 - HALO), r_hi = min(H, h0 + R + HALO);
    const int rows = min(R, H - h0);
    const uint32_t wbytes = (uint32_t)((r_hi - r_lo) * W * sizeof(T));
    const uint32_t cbytes = (uint32_t)(rows * W * sizeof(T));
    const int s = it % BWD_NSTAGE;
    T* sx = reinterpret_cast<T*>(smem_raw + (size_t)s * stage_bytes);
    T* sg = sx + BWD_CC * plane;
    T* sp = sg + BWD_CC * plane;
    const int woff = (r_lo - (h0 - HALO)) * W;
    const int c0 = ch * BWD_CC;
    const T* xsrc = (t ? x2 : x1) + b * sB + (int64_t)c0 * sC + r_lo * W;
    const T* gsrc = gout + ((int64_t)(b * C + c0) * 4) * HW;
    mbar_expect_tx(full + s, (2 * wbytes + cbytes) * BWD_CC);
#pragma unroll
    for (int cc = 0; cc < BWD_CC; ++cc) {
      bulk_g2s(sx + cc * plane + woff, xsrc + cc * sC, wbytes, full + s);
      bulk_g2s(sg + cc * plane + woff, gsrc + ((int64_t)cc * 4 + 1 + t) * HW + r_lo * W, wbytes, full + s);
      bulk_g2s(sp + cc * core, gsrc + ((int64_t)cc * 4 + (t ? 3 : 0)) * HW + h0 * W, cbytes, full + s);
    }
  };
  if (tid == 0) {
    const int pre = total_it < BWD_NSTAGE ? total_it : BWD_NSTAGE;
    for (int it = 0; it < pre; ++it) issue(it);
  }

  // ---- per-tile register state ----
  float lw[BWD_NP][BWD_K];      // gather list: weights
  int lo[BWD_NP][BWD_K];        // gather list: source offsets inside a window plane
  int ln[BWD_NP];               // hits (may exceed BWD_K: the rest is re-probed per chunk)
  int tpix[BWD_NP];             // own pixel h*W+w, or -1
  // own pixel as a SOURCE (flow gradient)
  float wx0[BWD_NP], wx1[BWD_NP], wy0[BWD_NP], wy1[BWD_NP], gix[BWD_NP], giy[BWD_NP], gate_x[BWD_NP], gate_y[BWD_NP];
  int soff[BWD_NP], gofs[BWD_NP];   // nw tap: window offset (or -1) and global offset*4 + flags
  int tb = 0, tt = 0, th0 = 0;
  int dxlo = 0, dxhi = 0, dylo = 0, dyhi = 0, wlo = 0, whi = 0;

  // visit every source of target (ty,tx) in fixed order: f(weight, window offset, hit index)
  auto probe = [&](int ty, int tx, auto&& f) {
    int n = 0;
    const int sy_a = max(wlo, ty - 1 - dyhi), sy_b = min(whi - 1, ty - dylo);
    const int sx_a = max(0, tx - 1 - dxhi), sx_b = min(W - 1, tx - dxlo);
    for (int sy = sy_a; sy <= sy_b; ++sy) {
      const int rowo = (sy - (th0 - HALO)) * W;
      for (int sx = sx_a; sx <= sx_b; ++sx) {
        const float ix = s_ix[rowo + sx], iy = s_iy[rowo + sx];
        const float x0f = floorf(ix), y0f = floorf(iy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const int ex = tx - x0, ey = ty - y0;              // 0: nw/sw column or nw/ne row; 1: the other
        if ((unsigned)ex > 1u || (unsigned)ey > 1u) continue;
        if (!near_col(x0 - sx, geo)) continue;
        const float wx = ex ? __fsub_rn(ix, x0f) : __fsub_rn(__fadd_rn(x0f, 1.f), ix);
        const float wy = ey ? __fsub_rn(iy, y0f) : __fsub_rn(__fadd_rn(y0f, 1.f), iy);
        f(__fmul_rn(wx, wy), rowo + sx, n);
        ++n;
      }
    }
    return n;
  };

  for (int it = 0; it < total_it; ++it) {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    if (ch == 0) {
      // ================= phase 0: window coordinates =================
      const int tile = blockIdx.x + tl * gridDim.x;
      const int band = tile % nbands, bt = tile / nbands;
      tb = bt >> 1; tt = bt & 1; th0 = band * R;
      wlo = max(0, th0 - HALO); whi = min(H, th0 + R + HALO);
      const float* fl = flow + ((int64_t)(tb * 2) * 2 + tt) * HW;
      if (tid < 4) s_rng[tid] = (tid & 1) ? -(1 << 30) : (1 << 30);
      __syncthreads();   // also orders the previous tile's probes (overflow path) before s_ix is rewritten
      int mn_x = 1 << 30, mx_x = -(1 << 30), mn_y = 1 << 30, mx_y = -(1 << 30);
      for (int i = tid + (wlo - (th0 - HALO)) * W; i < (whi - (th0 - HALO)) * W; i += BWD_THREADS) {
        const int r = i / W, col = i - r * W;
        const int sy = th0 - HALO + r;
        const int p = sy * W + col;
        const Axis ax = axis_coord(__ldg(xs + col), __ldg(fl + p), W);
        const Axis ay = axis_coord(__ldg(ys + sy), __ldg(fl + p + 2 * (int64_t)HW), H);
        s_ix[i] = ax.i; s_iy[i] = ay.i;
        const int dx = ax.i0 - col, dy = ay.i0 - sy;
        mn_x = min(mn_x, dx); mx_x = max(mx_x, dx); mn_y = min(mn_y, dy); mx_y = max(mx_y, dy);
      }
      mn_x = __reduce_min_sync(0xffffffffu, mn_x); mx_x = __reduce_max_sync(0xffffffffu, mx_x);
      mn_y = __reduce_min_sync(0xffffffffu, mn_y); mx_y = __reduce_max_sync(0xffffffffu, mx_y);
      if ((tid & 31) == 0) {
        atomicMin(s_rng + 0, mn_x); atomicMax(s_rng + 1, mx_x);
        atomicMin(s_rng + 2, mn_y); atomicMax(s_rng + 3, mx_y);
      }
      __syncthreads();
      dxlo = max(s_rng[0], -geo.DCAP); dxhi = min(s_rng[1], geo.DCAP);
      dylo = s_rng[2]; dyhi = s_rng[3];
      // ================= phase 1: gather lists + own footprints =================
#pragma unroll
      for (int k = 0; k < BWD_NP; ++k) {
        const int pl = tid + k * BWD_THREADS;
        const int r = pl / W, col = pl - r * W;
        const int h = th0 + r;
        gix[k] = giy[k] = 0.f;
#pragma unroll
        for (int j = 0; j < BWD_K; ++j) { lw[k][j] = 0.f; lo[k][j] = 0; }
        if (r < R && h < H) {
          tpix[k] = h * W + col;
          ln[k] = probe(h, col, [&](float w, int off, int n) {
#pragma unroll
            for (int j = 0; j < BWD_K; ++j)
              if (n == j) { lw[k][j] = w; lo[k][j] = off; }
          });
          const int p = tpix[k];
          const Footprint fp = footprint(__ldg(xs + col), __ldg(ys + h), __ldg(fl + p),
                                         __ldg(fl + p + 2 * (int64_t)HW), W, H);
          wx0[k] = fp.wx0; wx1[k] = fp.wx1; wy0[k] = fp.wy0; wy1[k] = fp.wy1;
          gate_x[k] = fp.gx_gate; gate_y[k] = fp.gy_gate;
          gofs[k] = (fp.y0 * W + fp.x0) * 4 + (fp.x1ok ? 1 : 0) + (fp.y1ok ? 2 : 0);
          const int sr = fp.y0 - (th0 - HALO);
          soff[k] = (sr >= 0 && sr + (fp.y1ok ? 1 : 0) < WR) ? sr * W + fp.x0 : -1;
        } else {
          tpix[k] = -1; ln[k] = 0; soff[k] = -1; gofs[k] = 0;
          wx0[k] = wx1[k] = wy0[k] = wy1[k] = gate_x[k] = gate_y[k] = 0.f;
        }
      }
    }
    // ================= phase 2: one channel chunk =================
    const int s = it % BWD_NSTAGE;
    mbar_wait(full + s, (uint32_t)((it / BWD_NSTAGE) & 1));
    const T* sx = reinterpret_cast<const T*>(smem_raw + (size_t)s * stage_bytes);
    const T* sg = sx + BWD_CC * plane;
    const T* sp = sg + BWD_CC * plane;
    const int c0 = ch * BWD_CC;
    T* gxo = (tt ? gx2 : gx1) + tb * sB + (int64_t)c0 * sC;
    const T* xg = (tt ? x2 : x1) + tb * sB + (int64_t)c0 * sC;
#pragma unroll
    for (int k = 0; k < BWD_NP; ++k) {
      if (tpix[k] < 0) continue;
      const int pl = tid + k * BWD_THREADS;      // == r*W + col inside the core buffer
      const int own = pl + HALO * W;             // own pixel inside a window plane
      const bool x1ok = gofs[k] & 1, y1ok = gofs[k] & 2;
#pragma unroll
      for (int cc = 0; cc < BWD_CC; ++cc) {
        const T* g = sg + cc * plane;
        // ---- target side: gx = pass-through + gathered scatter ----
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < BWD_K; ++j)
          if (j < ln[k]) acc = fmaf(lw[k][j], cvtf<T>(g[lo[k][j]]), acc);
        if (ln[k] > BWD_K) {
          const int h = tpix[k] / W;
          probe(h, tpix[k] - h * W, [&](float w, int off, int n) {
            if (n >= BWD_K) acc = fmaf(w, cvtf<T>(g[off]), acc);
          });
        }
        gxo[cc * sC + tpix[k]] = fromf<T>(__fadd_rn(cvtf<T>(sp[cc * core + pl]), acc));
        // ---- source side: flow gradient sums (ATen order nw, ne, sw, se; channels ascending) ----
        const float go = cvtf<T>(g[own]);
        float v_nw, v_ne = 0.f, v_sw = 0.f, v_se = 0.f;
        if (soff[k] >= 0) {
          const T* q = sx + cc * plane + soff[k];
          v_nw = cvtf<T>(q[0]);
          if (x1ok) v_ne = cvtf<T>(q[1]);
          if (y1ok) v_sw = cvtf<T>(q[W]);
          if (x1ok && y1ok) v_se = cvtf<T>(q[W + 1]);
        } else {
          const T* q = xg + cc * sC + (gofs[k] >> 2);
          v_nw = ldf(q);
          if (x1ok) v_ne = ldf(q + 1);
          if (y1ok) v_sw = ldf(q + W);
          if (x1ok && y1ok) v_se = ldf(q + W + 1);
        }
        gix[k] = fmaf(-__fmul_rn(v_nw, wy0[k]), go, gix[k]);
        giy[k] = fmaf(-__fmul_rn(v_nw, wx0[k]), go, giy[k]);
        if (x1ok) {
          gix[k] = fmaf(__fmul_rn(v_ne, wy0[k]), go, gix[k]);
          giy[k] = fmaf(-__fmul_rn(v_ne, wx1[k]), go, giy[k]);
        }
        if (y1ok) {
          gix[k] = fmaf(-__fmul_rn(v_sw, wy1[k]), go, gix[k]);
          giy[k] = fmaf(__fmul_rn(v_sw, wx0[k]), go, giy[k]);
        }
        if (x1ok && y1ok) {
          gix[k] = fmaf(__fmul_rn(v_se, wy1[k]), go, gix[k]);
          giy[k] = fmaf(__fmul_rn(v_se, wx1[k]), go, giy[k]);
        }
      }
    }
    if (ch == nchunk - 1) {   // ================= phase 3: flow gradient of this tile =================
#pragma unroll
      for (int k = 0; k < BWD_NP; ++k) {
        if (tpix[k] < 0) continue;
        const int64_t fo = ((int64_t)(tb * 2) * 2 + tt) * HW + tpix[k];
        const float mx = __fmul_rn(gate_x[k], __fmul_rn((float)(W - 1), 0.5f));
        const float my = __fmul_rn(gate_y[k], __fmul_rn((float)(H - 1), 0.5f));
        gflow[fo] = __fdiv_rn(__fmul_rn(mx, gix[k]), (float)W);
        gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, giy[k]), (float)H);
      }
    }
    __syncthreads();   // stage s fully consumed
    if (tid == 0 && it + BWD_NSTAGE < total_it) issue(it + BWD_NSTAGE);
  }
}

// Contributions the tile kernel does not own (source row outside the target tile's window, or column
// displacement beyond DCAP): same predicate, evaluated from the flow alone; L2 atomics.  Launched after
// the tile kernel on the same stream.
template <typename T> __device__ __forceinline__ void far_add(T* p, float v);
template <> __device__ __forceinline__ void far_add<float>(float* p, float v) { atomicAdd(p, v); }
template <> __device__ __forceinline__ void far_add<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

template <typename T>
__global__ void __launch_bounds__(256)
warp_bwd_far_kernel(const T* __restrict__ gout, const float* __restrict__ flow, const float* __restrict__ xs,
                    const float* __restrict__ ys, T* __restrict__ gx1, T* __restrict__ gx2, int64_t sB,
                    int64_t sC, int C, int H, int W, BwdGeom geo) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  const int h = p / W, w = p - h * W;
  const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
  const Footprint fp = footprint(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
  const bool far_c = !near_col(fp.x0 - w, geo);
  const bool far0 = far_c || !near_row(h, fp.y0, geo);
  const bool far1 = fp.y1ok && (far_c || !near_row(h, fp.y0 + 1, geo));
  if (!far0 && !far1) return;
  const float w00 = __fmul_rn(fp.wx0, fp.wy0), w01 = __fmul_rn(fp.wx1, fp.wy0);
  const float w10 = __fmul_rn(fp.wx0, fp.wy1), w11 = __fmul_rn(fp.wx1, fp.wy1);
  T* dst = (t ? gx2 : gx1) + b * sB + fp.y0 * W + fp.x0;
  const T* g = gout + ((int64_t)b * C * 4 + (1 + t)) * HW + p;
  for (int c = 0; c < C; ++c) {
    const float go = cvtf<T>(g[(int64_t)c * 4 * HW]);
    T* q = dst + c * sC;
    if (far0) {
      far_add(q, __fmul_rn(w00, go));
      if (fp.x1ok) far_add(q + 1, __fmul_rn(w01, go));
    }
    if (far1) {
      far_add(q + W, __fmul_rn(w10, go));
      if (fp.x1ok) far_add(q + W + 1, __fmul_rn(w11, go));
    }
  }
}

template <typename T>
int warp_bwd_tiled(const T* gout, const T* x1, const T* x2, int64_t sB, int64_t sC, const float* flow,
                   const float* xs, const float* ys, T* gx1, T* gx2, float* gflow, int B, int C, int H, int W,
                   cudaStream_t st) {
  const DeviceInfo di = device_info();
  BwdGeom geo;
  geo.HALO = option(OPT_BWD_HALO) < 1 ? 1 : option(OPT_BWD_HALO);
  geo.DCAP = geo.HALO + 1;
  int R = (BWD_THREADS * BWD_NP) / W;      // 1024 target pixels per tile
  if (R < 1) R = 1;
  if (R > H) R = H;
  if (R * W > BWD_THREADS * BWD_NP) return fail(SMOW_ERANGE, "W=%d too wide for the tiled backward", W);
  geo.R = R;
  const int WR = R + 2 * geo.HALO;
  const size_t stage = (size_t)BWD_CC * (2 * WR * W + R * W) * sizeof(T);
  const size_t smem = BWD_NSTAGE * stage + 2 * (size_t)WR * W * sizeof(float) + 4 * sizeof(int) +
                      BWD_NSTAGE * sizeof(uint64_t);
  if (smem > (size_t)di.smem_optin) return fail(SMOW_ERANGE, "tile does not fit shared memory (%zu B)", smem);
  auto kern = warp_bwd_tiled_kernel<T>;
  static thread_local size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  const int nbands = (H + R - 1) / R;
  const int ntiles = 2 * B * nbands;
  const int grid = ntiles < di.sms ? ntiles : di.sms;
  kern<<<grid, BWD_THREADS, smem, st>>>(gout, x1, x2, sB, sC, flow, xs, ys, gx1, gx2, gflow, C, H, W, geo, nbands,
                                        ntiles);
  warp_bwd_far_kernel<T><<<dim3((H * W + 255) / 256, 2 * B), 256, 0, st>>>(gout, flow, xs, ys, gx1, gx2, sB, sC, C,
                                                                           H, W, geo);
  count_launch(2);
  return check_launch("warp_bwd_tiled");
}

template int warp_bwd_tiled<float>(const float*, const float*, const float*, int64_t, int64_t, const float*,
                                   const float*, const float*, float*, float*, float*, int, int, int, int,
                                   cudaStream_t);
template int warp_bwd_tiled<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*,
                                           int64_t, int64_t, const float*, const float*, const float*,
                                           __nv_bfloat16*, __nv_bfloat16*, float*, int, int, int, int,
                                           cudaStream_t);
}  // namespace smow
