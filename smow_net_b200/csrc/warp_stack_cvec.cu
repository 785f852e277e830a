// A1, channel-vectorised tile kernels (variant 2, fp32 NCDHW in / NCDHW out, sm_100a).
//
// ncu on the plane-staged kernels (profiles/r1_*) showed them ISSUE-bound, not HBM-bound: one
// LDS.32 + one FFMA per tap per channel.  Here a tile is staged PIXEL-MAJOR in shared memory —
// 16 bytes (4 channels) per pixel, rows swizzled — so that one LDS.128 fetches a tap for four
// channels and one FFMA2 (Blackwell's packed fp32x2 FMA, scalar weight broadcast) does the
// multiply-add for two of them.  Global memory stays channel-planar (what cuDNN produces): the
// fill reads 4 pixels x 4 channels with four coalesced LDG.128, transposes in registers and
// writes four conflict-free STS.128; results go back with coalesced 128 B stores per channel.
//
// Software pipeline: global loads of chunk i+1 are issued before the math of chunk i and land
// in the other shared buffer after it (one __syncthreads per chunk, no mbarrier needed).
//
// Tiling, the inverse-gather backward, the register gather lists, the near/far predicate and
// the far-contribution side kernel are those of warp_stack_bwd_tiled.cu.
#include "warp_stack_tiled.cuh"
#include "bulk.cuh"

namespace smow {

constexpr int CV_THREADS = 512;
constexpr int CV_NP = 2;   // pixels per thread
constexpr int CV_K = 6;    // register gather-list length (bilinear scatter: 4 sources per target on average)
constexpr int CV_PF = 3;   // L2 prefetch distance, in channel chunks

struct CvGeom { int R, HALO, DCAP, WR, U, nbands, ntiles; };

// 16-byte slot swizzle: keeps both the transposed fill (lanes 4 pixels apart) and the tap reads
// (lanes 1 pixel apart) free of bank conflicts.  Only the low 3 bits change.
__device__ __forceinline__ int swz(int p) { return p ^ ((p >> 3) & 7); }

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, f2(-b.x, -b.y)); }

struct Unit { float4 c[4]; };   // 4 consecutive pixels of 4 channels

// unit u = window pixels 4u..4u+3 (same row since W % 4 == 0); plane0 points at channel c0 of the frame
__device__ __forceinline__ bool load_unit(Unit& v, const float* __restrict__ plane0, int64_t sC, int u,
                                          const CvGeom& g, int W, int H, int wr0) {
  const int p = 4 * u;
  const int row = p / W, col = p - row * W;
  const int y = wr0 + row;
  if (u >= g.U || y < 0 || y >= H) return false;
  const float* q = plane0 + (int64_t)y * W + col;
#pragma unroll
  for (int c = 0; c < 4; ++c) v.c[c] = __ldg(reinterpret_cast<const float4*>(q + c * sC));
  return true;
}
__device__ __forceinline__ void store_unit(float4* buf, int u, const Unit& v) {
  const int p = 4 * u, s = (p >> 3) & 7;
  buf[(p + 0) ^ s] = make_float4(v.c[0].x, v.c[1].x, v.c[2].x, v.c[3].x);
  buf[(p + 1) ^ s] = make_float4(v.c[0].y, v.c[1].y, v.c[2].y, v.c[3].y);
  buf[(p + 2) ^ s] = make_float4(v.c[0].z, v.c[1].z, v.c[2].z, v.c[3].z);
  buf[(p + 3) ^ s] = make_float4(v.c[0].w, v.c[1].w, v.c[2].w, v.c[3].w);
}

struct TileId { int b, t, h0; };
__device__ __forceinline__ TileId tile_of(int tile, const CvGeom& g) {
  const int band = tile % g.nbands, bt = tile / g.nbands;
  return TileId{bt >> 1, bt & 1, band * g.R};
}

// ------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_THREADS, 1)
warp_fwd_cvec_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int64_t sB, int64_t sC,
                     const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                     float* __restrict__ out, int C, int H, int W, CvGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int slots = g.WR * W;                               // pixels per window buffer (multiple of 8)
  float4* bufs[2] = {reinterpret_cast<float4*>(smem_raw), reinterpret_cast<float4*>(smem_raw) + slots};
  const int tid = threadIdx.x;
  const int nchunk = C >> 2;
  const int my_tiles = (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_it = my_tiles * nchunk;

  auto chunk_src = [&](int it, int& wr0) -> const float* {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    const TileId ti = tile_of(blockIdx.x + tl * gridDim.x, g);
    wr0 = ti.h0 - g.HALO;
    return (ti.t ? x2 : x1) + ti.b * sB + (int64_t)(ch * 4) * sC;
  };
  // L2 prefetch of pipeline iteration `it` (lanes 0..3 of warp 0: one channel plane each) and, at a
  // tile boundary, of the tile's flow rows: DRAM latency is taken off the register-staged fill
  auto prefetch = [&](int it) {
    if (it >= total_it || tid >= 8) return;
    const int tl = it / nchunk, ch = it - tl * nchunk;
    const TileId tj = tile_of(blockIdx.x + tl * gridDim.x, g);
    const int r_lo = max(0, tj.h0 - g.HALO), r_hi = min(H, tj.h0 + g.R + g.HALO);
    if (tid < 4) {
      const float* p = (tj.t ? x2 : x1) + tj.b * sB + (int64_t)(ch * 4 + tid) * sC + r_lo * W;
      bulk_prefetch_l2(p, (uint32_t)((r_hi - r_lo) * W * sizeof(float)));
    } else if (ch == 0 && tid < 6) {
      const int h1 = min(H, tj.h0 + g.R);
      const float* p = flow + ((int64_t)(tj.b * 2 + (tid - 4)) * 2 + tj.t) * HW + tj.h0 * W;
      bulk_prefetch_l2(p, (uint32_t)((h1 - tj.h0) * W * sizeof(float)));
    }
  };
  for (int d = 1; d <= CV_PF; ++d) prefetch(d);
  {  // prologue: chunk 0 -> buffer 0
    Unit v;
    int wr0;
    const float* src = chunk_src(0, wr0);
    if (total_it > 0 && load_unit(v, src, sC, tid, g, W, H, wr0)) store_unit(bufs[0], tid, v);
  }
  __syncthreads();

  float w_nw[CV_NP], w_ne[CV_NP], w_sw[CV_NP], w_se[CV_NP];
  int s_nw[CV_NP], s_ne[CV_NP], s_sw[CV_NP], s_se[CV_NP];   // swizzled slots, or s_nw = -1: global fallback
  int goff[CV_NP], opix[CV_NP];
  TileId ti{0, 0, 0};

  for (int it = 0; it < total_it; ++it) {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    if (ch == 0) {
      ti = tile_of(blockIdx.x + tl * gridDim.x, g);
      const float* fl = flow + ((int64_t)(ti.b * 2) * 2 + ti.t) * HW;
#pragma unroll
      for (int k = 0; k < CV_NP; ++k) {
        const int pl = tid + k * CV_THREADS;
        const int r = pl / W, col = pl - r * W, h = ti.h0 + r;
        if (r < g.R && h < H) {
          const int p = h * W + col;
          const Footprint fp = footprint(__ldg(xs + col), __ldg(ys + h), __ldg(fl + p),
                                         __ldg(fl + p + 2 * (int64_t)HW), W, H);
          w_nw[k] = __fmul_rn(fp.wx0, fp.wy0);
          w_ne[k] = fp.x1ok ? __fmul_rn(fp.wx1, fp.wy0) : 0.f;
          w_sw[k] = fp.y1ok ? __fmul_rn(fp.wx0, fp.wy1) : 0.f;
          w_se[k] = (fp.x1ok && fp.y1ok) ? __fmul_rn(fp.wx1, fp.wy1) : 0.f;
          const int dx = fp.x1ok ? 1 : 0, dy = fp.y1ok ? W : 0;
          goff[k] = (fp.y0 * W + fp.x0) * 4 + (fp.x1ok ? 1 : 0) + (fp.y1ok ? 2 : 0);
          const int sr = fp.y0 - (ti.h0 - g.HALO);
          if (sr >= 0 && sr + (fp.y1ok ? 1 : 0) < g.WR) {
            const int q = sr * W + fp.x0;
            s_nw[k] = swz(q); s_ne[k] = swz(q + dx); s_sw[k] = swz(q + dy); s_se[k] = swz(q + dy + dx);
          } else {
            s_nw[k] = -1; s_ne[k] = s_sw[k] = s_se[k] = 0;
          }
          opix[k] = p;
        } else {
          opix[k] = -1; s_nw[k] = -1; s_ne[k] = s_sw[k] = s_se[k] = 0; goff[k] = 0;
          w_nw[k] = w_ne[k] = w_sw[k] = w_se[k] = 0.f;
        }
      }
    }
    prefetch(it + 1 + CV_PF);
    // stage the next chunk in registers
    Unit nxt;
    bool have_nxt = false;
    if (it + 1 < total_it) {
      int wr0;
      const float* src = chunk_src(it + 1, wr0);
      have_nxt = load_unit(nxt, src, sC, tid, g, W, H, wr0);
    }
    const float4* X = bufs[it & 1];
    const int c0 = ch * 4;
    float* ob = out + ((int64_t)(ti.b * C + c0) * 4) * HW;
    const float* gsrc = (ti.t ? x2 : x1) + ti.b * sB + (int64_t)c0 * sC;
#pragma unroll
    for (int k = 0; k < CV_NP; ++k) {
      if (opix[k] < 0) continue;
      float4 r;
      if (s_nw[k] >= 0) {
        const float4 a = X[s_nw[k]], b = X[s_ne[k]], c = X[s_sw[k]], d = X[s_se[k]];
        // ATen order nw, ne, sw, se; invalid taps carry weight 0 and a clamped (valid) slot
        float2 lo = __fmul2_rn(bc(w_nw[k]), f2(a.x, a.y)), hi = __fmul2_rn(bc(w_nw[k]), f2(a.z, a.w));
        lo = __ffma2_rn(bc(w_ne[k]), f2(b.x, b.y), lo); hi = __ffma2_rn(bc(w_ne[k]), f2(b.z, b.w), hi);
        lo = __ffma2_rn(bc(w_sw[k]), f2(c.x, c.y), lo); hi = __ffma2_rn(bc(w_sw[k]), f2(c.z, c.w), hi);
        lo = __ffma2_rn(bc(w_se[k]), f2(d.x, d.y), lo); hi = __ffma2_rn(bc(w_se[k]), f2(d.z, d.w), hi);
        r = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {  // footprint outside the staged rows
        const bool x1ok = goff[k] & 1, y1ok = goff[k] & 2;
        const int o = goff[k] >> 2, dx = x1ok ? 1 : 0, dy = y1ok ? W : 0;
        float t[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* q = gsrc + c * sC + o;
          float acc = __fmul_rn(__ldg(q), w_nw[k]);
          acc = fmaf(__ldg(q + dx), w_ne[k], acc);
          acc = fmaf(__ldg(q + dy), w_sw[k], acc);
          acc = fmaf(__ldg(q + dy + dx), w_se[k], acc);
          t[c] = acc;
        }
        r = make_float4(t[0], t[1], t[2], t[3]);
      }
      const float4 own = X[swz(tid + k * CV_THREADS + g.HALO * W)];   // un-warped slot
      float* ow = ob + (int64_t)(1 + ti.t) * HW + opix[k];
      float* op = ob + (int64_t)(ti.t ? 3 : 0) * HW + opix[k];
      const int64_t cs = (int64_t)4 * HW;
      ow[0] = r.x; ow[cs] = r.y; ow[2 * cs] = r.z; ow[3 * cs] = r.w;
      op[0] = own.x; op[cs] = own.y; op[2 * cs] = own.z; op[3 * cs] = own.w;
    }
    if (have_nxt) store_unit(bufs[(it + 1) & 1], tid, nxt);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------
__device__ __forceinline__ bool cv_near_row(int sy, int ty, const CvGeom& g) {
  const int h0 = (ty / g.R) * g.R;
  return sy >= h0 - g.HALO && sy < h0 + g.R + g.HALO;
}
__device__ __forceinline__ bool cv_near_col(int dx, const CvGeom& g) { return dx >= -g.DCAP && dx <= g.DCAP; }

__global__ void __launch_bounds__(CV_THREADS, 1)
warp_bwd_cvec_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                     int64_t sB, int64_t sC, const float* __restrict__ flow, const float* __restrict__ xs,
                     const float* __restrict__ ys, float* __restrict__ gx1, float* __restrict__ gx2,
                     float* __restrict__ gflow, int C, int H, int W, CvGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int slots = g.WR * W;
  float4* base4 = reinterpret_cast<float4*>(smem_raw);
  // [buffer 0: x window | gout window][buffer 1: x | gout][ix][iy][range]
  float* s_ix = reinterpret_cast<float*>(base4 + 4 * slots);
  float* s_iy = s_ix + slots;
  int* s_rng = reinterpret_cast<int*>(s_iy + slots);
  const int tid = threadIdx.x;
  const int nchunk = C >> 2;
  const int my_tiles = (g.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_it = my_tiles * nchunk;

  auto fetch = [&](int it, Unit& vx, Unit& vg, bool& ok) {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    const TileId ti = tile_of(blockIdx.x + tl * gridDim.x, g);
    const int c0 = ch * 4;
    const float* xsrc = (ti.t ? x2 : x1) + ti.b * sB + (int64_t)c0 * sC;
    const float* gsrc = gout + ((int64_t)(ti.b * C + c0) * 4 + 1 + ti.t) * HW;
    ok = load_unit(vx, xsrc, sC, tid, g, W, H, ti.h0 - g.HALO);
    if (ok) load_unit(vg, gsrc, (int64_t)4 * HW, tid, g, W, H, ti.h0 - g.HALO);
  };
  auto prefetch = [&](int it) {   // lanes 0..11: x window, gout window, gout pass rows; 12..13: flow rows
    if (it >= total_it || tid >= 14) return;
    const int tl = it / nchunk, ch = it - tl * nchunk;
    const TileId tj = tile_of(blockIdx.x + tl * gridDim.x, g);
    const int r_lo = max(0, tj.h0 - g.HALO), r_hi = min(H, tj.h0 + g.R + g.HALO);
    const int h1 = min(H, tj.h0 + g.R);
    const uint32_t wbytes = (uint32_t)((r_hi - r_lo) * W * sizeof(float));
    const int c = ch * 4 + (tid & 3);
    if (tid < 4) {
      bulk_prefetch_l2((tj.t ? x2 : x1) + tj.b * sB + (int64_t)c * sC + r_lo * W, wbytes);
    } else if (tid < 8) {
      bulk_prefetch_l2(gout + ((int64_t)(tj.b * C + c) * 4 + 1 + tj.t) * HW + r_lo * W, wbytes);
    } else if (tid < 12) {
      bulk_prefetch_l2(gout + ((int64_t)(tj.b * C + c) * 4 + (tj.t ? 3 : 0)) * HW + tj.h0 * W,
                       (uint32_t)((h1 - tj.h0) * W * sizeof(float)));
    } else if (ch == 0) {
      bulk_prefetch_l2(flow + ((int64_t)(tj.b * 2 + (tid - 12)) * 2 + tj.t) * HW + r_lo * W, wbytes);
    }
  };
  for (int d = 1; d <= CV_PF; ++d) prefetch(d);
  {
    Unit vx, vg;
    bool ok = false;
    if (total_it > 0) fetch(0, vx, vg, ok);
    if (ok) { store_unit(base4, tid, vx); store_unit(base4 + slots, tid, vg); }
  }
  __syncthreads();

  float lw[CV_NP][CV_K];
  int lo[CV_NP][CV_K];          // swizzled slots of the gather sources
  int ln[CV_NP], tpix[CV_NP];
  float wx0[CV_NP], wx1[CV_NP], wy0[CV_NP], wy1[CV_NP], gate_x[CV_NP], gate_y[CV_NP];
  float2 gixp[CV_NP], giyp[CV_NP];
  int s_nw[CV_NP], s_ne[CV_NP], s_sw[CV_NP], s_se[CV_NP], gofs[CV_NP];
  TileId ti{0, 0, 0};
  int dxlo = 0, dxhi = 0, dylo = 0, dyhi = 0, wlo = 0, whi = 0;

  auto probe = [&](int ty, int tx, auto&& f) {
    int n = 0;
    const int sy_a = max(wlo, ty - 1 - dyhi), sy_b = min(whi - 1, ty - dylo);
    const int sx_a = max(0, tx - 1 - dxhi), sx_b = min(W - 1, tx - dxlo);
    for (int sy = sy_a; sy <= sy_b; ++sy) {
      const int rowo = (sy - (ti.h0 - g.HALO)) * W;
      for (int sx = sx_a; sx <= sx_b; ++sx) {
        const float ix = s_ix[rowo + sx], iy = s_iy[rowo + sx];
        const float x0f = floorf(ix), y0f = floorf(iy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const int ex = tx - x0, ey = ty - y0;
        if ((unsigned)ex > 1u || (unsigned)ey > 1u) continue;
        if (!cv_near_col(x0 - sx, g)) continue;
        const float wx = ex ? __fsub_rn(ix, x0f) : __fsub_rn(__fadd_rn(x0f, 1.f), ix);
        const float wy = ey ? __fsub_rn(iy, y0f) : __fsub_rn(__fadd_rn(y0f, 1.f), iy);
        f(__fmul_rn(wx, wy), swz(rowo + sx), n);
        ++n;
      }
    }
    return n;
  };

  for (int it = 0; it < total_it; ++it) {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    if (ch == 0) {
      // ---------------- phase 0: window sample coordinates + displacement range ----------------
      ti = tile_of(blockIdx.x + tl * gridDim.x, g);
      wlo = max(0, ti.h0 - g.HALO); whi = min(H, ti.h0 + g.R + g.HALO);
      const float* fl = flow + ((int64_t)(ti.b * 2) * 2 + ti.t) * HW;
      if (tid < 4) s_rng[tid] = (tid & 1) ? -(1 << 30) : (1 << 30);
      __syncthreads();
      int mn_x = 1 << 30, mx_x = -(1 << 30), mn_y = 1 << 30, mx_y = -(1 << 30);
      for (int i = tid + (wlo - (ti.h0 - g.HALO)) * W; i < (whi - (ti.h0 - g.HALO)) * W; i += CV_THREADS) {
        const int r = i / W, col = i - r * W;
        const int sy = ti.h0 - g.HALO + r;
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
      dxlo = max(s_rng[0], -g.DCAP); dxhi = min(s_rng[1], g.DCAP);
      dylo = s_rng[2]; dyhi = s_rng[3];
      // ---------------- phase 1: gather lists (targets) and footprints (sources) ----------------
#pragma unroll
      for (int k = 0; k < CV_NP; ++k) {
        const int pl = tid + k * CV_THREADS;
        const int r = pl / W, col = pl - r * W, h = ti.h0 + r;
        gixp[k] = giyp[k] = f2(0.f, 0.f);
        const int own = swz(pl + g.HALO * W);
#pragma unroll
        for (int j = 0; j < CV_K; ++j) { lw[k][j] = 0.f; lo[k][j] = own; }
        if (r < g.R && h < H) {
          tpix[k] = h * W + col;
          ln[k] = probe(h, col, [&](float w, int slot, int n) {
#pragma unroll
            for (int j = 0; j < CV_K; ++j)
              if (n == j) { lw[k][j] = w; lo[k][j] = slot; }
          });
          const int p = tpix[k];
          const Footprint fp = footprint(__ldg(xs + col), __ldg(ys + h), __ldg(fl + p),
                                         __ldg(fl + p + 2 * (int64_t)HW), W, H);
          wx0[k] = fp.wx0; wx1[k] = fp.wx1; wy0[k] = fp.wy0; wy1[k] = fp.wy1;
          gate_x[k] = fp.gx_gate; gate_y[k] = fp.gy_gate;
          gofs[k] = (fp.y0 * W + fp.x0) * 4 + (fp.x1ok ? 1 : 0) + (fp.y1ok ? 2 : 0);
          const int sr = fp.y0 - (ti.h0 - g.HALO);
          const int dx = fp.x1ok ? 1 : 0, dy = fp.y1ok ? W : 0;
          if (sr >= 0 && sr + (fp.y1ok ? 1 : 0) < g.WR) {
            const int q = sr * W + fp.x0;
            s_nw[k] = swz(q); s_ne[k] = swz(q + dx); s_sw[k] = swz(q + dy); s_se[k] = swz(q + dy + dx);
          } else {
            s_nw[k] = -1; s_ne[k] = s_sw[k] = s_se[k] = 0;
          }
        } else {
          tpix[k] = -1; ln[k] = 0; s_nw[k] = -1; s_ne[k] = s_sw[k] = s_se[k] = 0; gofs[k] = 0;
          wx0[k] = wx1[k] = wy0[k] = wy1[k] = gate_x[k] = gate_y[k] = 0.f;
        }
      }
    }
    // ---------------- phase 2: one chunk of 4 channels ----------------
    prefetch(it + 1 + CV_PF);
    Unit nx, ng;
    bool have_nxt = false;
    if (it + 1 < total_it) fetch(it + 1, nx, ng, have_nxt);
    const float4* X = base4 + (size_t)(it & 1) * 2 * slots;
    const float4* G = X + slots;
    const int c0 = ch * 4;
    const int64_t cs = (int64_t)4 * HW;
    const float* gpass = gout + ((int64_t)(ti.b * C + c0) * 4 + (ti.t ? 3 : 0)) * HW;
    float* gxo = (ti.t ? gx2 : gx1) + ti.b * sB + (int64_t)c0 * sC;
    const float* xg = (ti.t ? x2 : x1) + ti.b * sB + (int64_t)c0 * sC;
#pragma unroll
    for (int k = 0; k < CV_NP; ++k) {
      if (tpix[k] < 0) continue;
      const float* gp = gpass + tpix[k];
      const float p0 = __ldg(gp), p1 = __ldg(gp + cs), p2 = __ldg(gp + 2 * cs), p3 = __ldg(gp + 3 * cs);
      // target side: gather the scatter (zero-weight padding instead of predicates)
      float2 a01 = f2(0.f, 0.f), a23 = f2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < CV_K; ++j) {
        const float4 v = G[lo[k][j]];
        a01 = __ffma2_rn(bc(lw[k][j]), f2(v.x, v.y), a01);
        a23 = __ffma2_rn(bc(lw[k][j]), f2(v.z, v.w), a23);
      }
      if (ln[k] > CV_K) {
        const int h = tpix[k] / W;
        probe(h, tpix[k] - h * W, [&](float w, int slot, int n) {
          if (n >= CV_K) {
            const float4 v = G[slot];
            a01 = __ffma2_rn(bc(w), f2(v.x, v.y), a01);
            a23 = __ffma2_rn(bc(w), f2(v.z, v.w), a23);
          }
        });
      }
      // source side: flow-gradient sums, factored:  gix += go * (wy0 (ne-nw) + wy1 (se-sw)),
      //                                             giy += go * (wx0 (sw-nw) + wx1 (se-ne))
      // (out-of-bounds taps alias an in-bounds slot; their weight wx1 / wy1 is exactly 0 there and
      // the other axis is gated off, so the sums equal ATen's skip-the-tap form)
      const float4 go = G[swz(tid + k * CV_THREADS + g.HALO * W)];
      float4 vnw, vne, vsw, vse;
      if (s_nw[k] >= 0) {
        vnw = X[s_nw[k]]; vne = X[s_ne[k]]; vsw = X[s_sw[k]]; vse = X[s_se[k]];
      } else {
        const bool x1ok = gofs[k] & 1, y1ok = gofs[k] & 2;
        const int o = gofs[k] >> 2, dx = x1ok ? 1 : 0, dy = y1ok ? W : 0;
        float t[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* q = xg + c * sC + o;
          t[0][c] = __ldg(q); t[1][c] = __ldg(q + dx); t[2][c] = __ldg(q + dy); t[3][c] = __ldg(q + dy + dx);
        }
        vnw = make_float4(t[0][0], t[0][1], t[0][2], t[0][3]); vne = make_float4(t[1][0], t[1][1], t[1][2], t[1][3]);
        vsw = make_float4(t[2][0], t[2][1], t[2][2], t[2][3]); vse = make_float4(t[3][0], t[3][1], t[3][2], t[3][3]);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float2 nw = half ? f2(vnw.z, vnw.w) : f2(vnw.x, vnw.y), ne = half ? f2(vne.z, vne.w) : f2(vne.x, vne.y);
        const float2 sw = half ? f2(vsw.z, vsw.w) : f2(vsw.x, vsw.y), se = half ? f2(vse.z, vse.w) : f2(vse.x, vse.y);
        const float2 gh = half ? f2(go.z, go.w) : f2(go.x, go.y);
        float2 tx = __fmul2_rn(bc(wy0[k]), sub2(ne, nw));
        tx = __ffma2_rn(bc(wy1[k]), sub2(se, sw), tx);
        float2 ty = __fmul2_rn(bc(wx0[k]), sub2(sw, nw));
        ty = __ffma2_rn(bc(wx1[k]), sub2(se, ne), ty);
        gixp[k] = __ffma2_rn(gh, tx, gixp[k]);
        giyp[k] = __ffma2_rn(gh, ty, giyp[k]);
      }
      float* o = gxo + tpix[k];
      o[0] = __fadd_rn(p0, a01.x); o[sC] = __fadd_rn(p1, a01.y);
      o[2 * sC] = __fadd_rn(p2, a23.x); o[3 * sC] = __fadd_rn(p3, a23.y);
    }
    if (ch == nchunk - 1) {
#pragma unroll
      for (int k = 0; k < CV_NP; ++k) {
        if (tpix[k] < 0) continue;
        const int64_t fo = ((int64_t)(ti.b * 2) * 2 + ti.t) * HW + tpix[k];
        const float mx = __fmul_rn(gate_x[k], __fmul_rn((float)(W - 1), 0.5f));
        const float my = __fmul_rn(gate_y[k], __fmul_rn((float)(H - 1), 0.5f));
        gflow[fo] = __fdiv_rn(__fmul_rn(mx, __fadd_rn(gixp[k].x, gixp[k].y)), (float)W);
        gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, __fadd_rn(giyp[k].x, giyp[k].y)), (float)H);
      }
    }
    if (have_nxt) {
      float4* nb = base4 + (size_t)((it + 1) & 1) * 2 * slots;
      store_unit(nb, tid, nx);
      store_unit(nb + slots, tid, ng);
    }
    __syncthreads();
  }
}

// far contributions: same predicate as the tile kernel, from the flow alone (see warp_stack_bwd_tiled.cu)
__global__ void __launch_bounds__(256)
warp_bwd_cvec_far_kernel(const float* __restrict__ gout, const float* __restrict__ flow,
                         const float* __restrict__ xs, const float* __restrict__ ys, float* __restrict__ gx1,
                         float* __restrict__ gx2, int64_t sB, int64_t sC, int C, int H, int W, CvGeom g) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const int b = blockIdx.y >> 1, t = blockIdx.y & 1;
  const int h = p / W, w = p - h * W;
  const int64_t fo = ((int64_t)(b * 2) * 2 + t) * HW + p;
  const Footprint fp = footprint(__ldg(xs + w), __ldg(ys + h), __ldg(flow + fo), __ldg(flow + fo + 2 * (int64_t)HW), W, H);
  const bool far_c = !cv_near_col(fp.x0 - w, g);
  const bool far0 = far_c || !cv_near_row(h, fp.y0, g);
  const bool far1 = fp.y1ok && (far_c || !cv_near_row(h, fp.y0 + 1, g));
  if (!far0 && !far1) return;
  const float w00 = __fmul_rn(fp.wx0, fp.wy0), w01 = __fmul_rn(fp.wx1, fp.wy0);
  const float w10 = __fmul_rn(fp.wx0, fp.wy1), w11 = __fmul_rn(fp.wx1, fp.wy1);
  float* dst = (t ? gx2 : gx1) + b * sB + fp.y0 * W + fp.x0;
  const float* gg = gout + ((int64_t)b * C * 4 + (1 + t)) * HW + p;
  for (int c = 0; c < C; ++c) {
    const float go = __ldg(gg + (int64_t)c * 4 * HW);
    float* q = dst + c * sC;
    if (far0) {
      atomicAdd(q, __fmul_rn(w00, go));
      if (fp.x1ok) atomicAdd(q + 1, __fmul_rn(w01, go));
    }
    if (far1) {
      atomicAdd(q + W, __fmul_rn(w10, go));
      if (fp.x1ok) atomicAdd(q + W + 1, __fmul_rn(w11, go));
    }
  }
}

// ------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------
static bool cv_geometry(CvGeom& g, int B, int H, int W, int halo) {
  g.HALO = halo < 1 ? 1 : halo;
  g.DCAP = g.HALO + 1;
  int R = (CV_THREADS * CV_NP) / W;
  if (R < 1) return false;
  if (R > H) R = H;
  g.R = R;
  g.WR = R + 2 * g.HALO;
  if ((g.WR * W) % 8 != 0) return false;
  g.U = g.WR * W / 4;
  if (g.U > CV_THREADS) return false;          // one fill unit per thread per array
  g.nbands = (H + R - 1) / R;
  g.ntiles = 2 * B * g.nbands;
  return true;
}

template <typename Kern>
static int cv_prepare(Kern kern, size_t smem, size_t& configured) {
  if (smem > (size_t)device_info().smem_optin) return SMOW_ERANGE;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  return 0;
}

int warp_fwd_cvec(const float* x1, const float* x2, int64_t sB, int64_t sC, const float* flow, const float* xs,
                  const float* ys, float* out, int B, int C, int H, int W, cudaStream_t st) {
  CvGeom g;
  if (!cv_geometry(g, B, H, W, option(OPT_FWD_HALO))) return SMOW_ERANGE;
  const size_t smem = 2 * (size_t)g.WR * W * sizeof(float4);
  static thread_local size_t configured = 0;
  if (int e = cv_prepare(warp_fwd_cvec_kernel, smem, configured)) return e;
  const int sms = device_info().sms;
  const int grid = g.ntiles < sms ? g.ntiles : sms;
  warp_fwd_cvec_kernel<<<grid, CV_THREADS, smem, st>>>(x1, x2, sB, sC, flow, xs, ys, out, C, H, W, g);
  count_launch();
  return check_launch("warp_fwd_cvec");
}

int warp_bwd_cvec(const float* gout, const float* x1, const float* x2, int64_t sB, int64_t sC, const float* flow,
                  const float* xs, const float* ys, float* gx1, float* gx2, float* gflow, int B, int C, int H,
                  int W, cudaStream_t st) {
  CvGeom g;
  if (!cv_geometry(g, B, H, W, option(OPT_BWD_HALO))) return SMOW_ERANGE;
  const size_t smem = 4 * (size_t)g.WR * W * sizeof(float4) + 2 * (size_t)g.WR * W * sizeof(float) + 4 * sizeof(int);
  static thread_local size_t configured = 0;
  if (int e = cv_prepare(warp_bwd_cvec_kernel, smem, configured)) return e;
  const int sms = device_info().sms;
  const int grid = g.ntiles < sms ? g.ntiles : sms;
  warp_bwd_cvec_kernel<<<grid, CV_THREADS, smem, st>>>(gout, x1, x2, sB, sC, flow, xs, ys, gx1, gx2, gflow, C, H, W, g);
  warp_bwd_cvec_far_kernel<<<dim3((H * W + 255) / 256, 2 * B), 256, 0, st>>>(gout, flow, xs, ys, gx1, gx2, sB, sC, C,
                                                                            H, W, g);
  count_launch(2);
  return check_launch("warp_bwd_cvec");
}

}  // namespace smow
