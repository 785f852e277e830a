// A1, channel-vectorised tile kernels (variant 2, fp32 NCDHW in / NCDHW out, sm_100a).
//
// ncu on the plane-staged kernels (profiles/r1_*) showed them ISSUE-bound, not HBM-bound: one
// LDS.32 + one FFMA per tap per channel.  Here a tile is staged PIXEL-MAJOR in shared memory —
// 16 bytes (4 channels) per pixel, slots swizzled — so that one LDS.128 fetches a tap for four
// channels and one FFMA2 (Blackwell's packed fp32x2 FMA, scalar weight broadcast) does the
// multiply-add for two of them.  Global memory stays channel-planar (what cuDNN produces): the
// fill reads 4 pixels x 4 channels with four coalesced LDG.128, transposes in registers and
// writes four conflict-free STS.128; results go back with coalesced 128 B stores per channel.
//
// Pipeline: a few lanes issue cp.async.bulk.prefetch.L2 for the chunk CV_PF steps ahead (no
// registers, no smem); the register-staged fill of chunk i+1 is issued before the math of chunk
// i and lands in the other shared buffer after it (one __syncthreads per chunk).  All index
// arithmetic is hoisted to tile granularity — the first version of this file spent most of its
// issue slots on integer divisions (profiles/r1_notes.md).
//
// Tiling, the inverse-gather backward, the register gather lists, the near/far predicate and
// the far-contribution side kernel are those of warp_stack_bwd_tiled.cu.
#include "warp_stack_tiled.cuh"
#include "bulk.cuh"

namespace smow {

constexpr int CV_THREADS = 256;
constexpr int CV_CTAS = 2;    // resident CTAs per SM the kernels are sized for
constexpr int CV_NP = 2;   // pixels per thread
constexpr int CV_K = 6;    // register gather-list length (bilinear scatter: 4 sources per target on average)
constexpr int CV_PF = 3;   // L2 prefetch distance, in channel chunks

struct CvGeom { int R, HALO, DCAP, WR, U, nbands, ntiles, PF; };

// 16-byte slot swizzle: keeps both the transposed fill (lanes 4 pixels apart) and the tap reads
// (lanes 1 pixel apart) free of bank conflicts.  Only the low 3 bits change.
__device__ __forceinline__ int swz(int p) { return p ^ ((p >> 3) & 7); }

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, f2(-b.x, -b.y)); }

struct Unit { float4 c[4]; };   // 4 consecutive pixels of 4 channels

__device__ __forceinline__ void load_unit(Unit& v, const float* __restrict__ q, int64_t sC) {
#pragma unroll
  for (int c = 0; c < 4; ++c) v.c[c] = __ldg(reinterpret_cast<const float4*>(q + c * sC));
}
// p4 = first window pixel of the unit (multiple of 4)
__device__ __forceinline__ void store_unit(float4* buf, int p4, const Unit& v) {
  const int s = (p4 >> 3) & 7;
  buf[(p4 + 0) ^ s] = make_float4(v.c[0].x, v.c[1].x, v.c[2].x, v.c[3].x);
  buf[(p4 + 1) ^ s] = make_float4(v.c[0].y, v.c[1].y, v.c[2].y, v.c[3].y);
  buf[(p4 + 2) ^ s] = make_float4(v.c[0].z, v.c[1].z, v.c[2].z, v.c[3].z);
  buf[(p4 + 3) ^ s] = make_float4(v.c[0].w, v.c[1].w, v.c[2].w, v.c[3].w);
}

struct TileId { int b, t, h0; };
__device__ __forceinline__ TileId tile_of(int tile, const CvGeom& g) {
  const int band = tile % g.nbands, bt = tile / g.nbands;
  return TileId{bt >> 1, bt & 1, band * g.R};
}

// ------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(CV_THREADS, CV_CTAS)
warp_fwd_cvec_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int64_t sB, int64_t sC,
                     const float* __restrict__ flow, const float* __restrict__ xs, const float* __restrict__ ys,
                     float* __restrict__ out, int C, int H, int W, CvGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int slots = g.WR * W;                    // pixels per window buffer (multiple of 8)
  float4* buf0 = reinterpret_cast<float4*>(smem_raw);
  const int tid = threadIdx.x;
  const int nchunk = C >> 2;
  const int64_t cs = (int64_t)4 * HW;            // channel stride of `out`
  // this thread's fill unit: window pixels 4*tid .. 4*tid+3 (one row, W % 4 == 0)
  const int u_p4 = 4 * tid;
  const int u_row = u_p4 / W, u_col = u_p4 - u_row * W;
  const bool u_any = tid < g.U;
  // this thread's output pixels inside a tile
  int p_r[CV_NP], p_c[CV_NP];
#pragma unroll
  for (int k = 0; k < CV_NP; ++k) {
    const int pl = tid + k * CV_THREADS;
    p_r[k] = pl / W; p_c[k] = pl - p_r[k] * W;
  }

  int tile = blockIdx.x;
  if (tile >= g.ntiles) return;
  TileId ti = tile_of(tile, g);
  // fill pointer of the chunk to be staged next, and whether this thread's unit row is inside the image
  auto unit_ptr = [&](const TileId& t, int ch) -> const float* {
    return (t.t ? x2 : x1) + t.b * sB + (int64_t)(ch * 4) * sC + (int64_t)(t.h0 - g.HALO + u_row) * W + u_col;
  };
  auto unit_ok = [&](const TileId& t) { const int y = t.h0 - g.HALO + u_row; return u_any && y >= 0 && y < H; };
  auto prefetch = [&](const TileId& t, int ch) {   // lanes 0..3: one channel plane each; 4..5: flow rows
    const int r_lo = max(0, t.h0 - g.HALO), r_hi = min(H, t.h0 + g.R + g.HALO);
    if (tid < 4) {
      bulk_prefetch_l2((t.t ? x2 : x1) + t.b * sB + (int64_t)(ch * 4 + tid) * sC + r_lo * W,
                       (uint32_t)((r_hi - r_lo) * W * sizeof(float)));
    } else if (ch == 0 && tid < 6) {
      bulk_prefetch_l2(flow + ((int64_t)(t.b * 2 + (tid - 4)) * 2 + t.t) * HW + t.h0 * W,
                       (uint32_t)((min(H, t.h0 + g.R) - t.h0) * W * sizeof(float)));
    }
  };
  {  // prologue: first chunk -> buffer 0
    if (tid < 8)
      for (int d = 1; d <= g.PF && d < nchunk; ++d) prefetch(ti, d);
    if (unit_ok(ti)) { Unit v; load_unit(v, unit_ptr(ti, 0), sC); store_unit(buf0, u_p4, v); }
  }
  __syncthreads();

  int parity = 0;
  for (; tile < g.ntiles; tile += gridDim.x) {
    const int ntile = tile + gridDim.x;
    const bool has_next_tile = ntile < g.ntiles;
    const TileId tn = has_next_tile ? tile_of(ntile, g) : ti;
    // ---- per-tile: footprints of this thread's pixels ----
    float w_nw[CV_NP], w_ne[CV_NP], w_sw[CV_NP], w_se[CV_NP];
    int s_nw[CV_NP], s_ne[CV_NP], s_sw[CV_NP], s_se[CV_NP];   // swizzled slots; s_nw = -1: global fallback
    int goff[CV_NP], s_own[CV_NP];
    float* o_warp[CV_NP];                                       // out pointer of channel 0, warped slot; 0 = masked
    const float* fl = flow + ((int64_t)(ti.b * 2) * 2 + ti.t) * HW;
#pragma unroll
    for (int k = 0; k < CV_NP; ++k) {
      const int h = ti.h0 + p_r[k];
      o_warp[k] = nullptr; s_nw[k] = -1; s_ne[k] = s_sw[k] = s_se[k] = 0; goff[k] = 0; s_own[k] = 0;
      w_nw[k] = w_ne[k] = w_sw[k] = w_se[k] = 0.f;
      if (p_r[k] < g.R && h < H) {
        const int p = h * W + p_c[k];
        const Footprint fp = footprint(__ldg(xs + p_c[k]), __ldg(ys + h), __ldg(fl + p), __ldg(fl + p + 2 * (int64_t)HW), W, H);
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
        }
        s_own[k] = swz((p_r[k] + g.HALO) * W + p_c[k]);
        o_warp[k] = out + ((int64_t)ti.b * C * 4 + (1 + ti.t)) * HW + p;
      }
    }
    const int64_t pass_delta = (int64_t)((ti.t ? 3 : 0) - (1 + ti.t)) * HW;   // warped slot -> un-warped slot
    const float* gsrc = (ti.t ? x2 : x1) + ti.b * sB;
    const bool ok_here = unit_ok(ti), ok_next = unit_ok(tn);
    const float* fill = unit_ptr(ti, 1);                                        // chunk ch+1 of this tile
    const float* fill_next_tile = unit_ptr(tn, 0);

    for (int ch = 0; ch < nchunk; ++ch) {
      // L2 prefetch CV_PF chunks ahead (this tile, else the next one)
      if (tid < 8) {
        const int pc = ch + 1 + g.PF;
        if (pc < nchunk) prefetch(ti, pc);
        else if (has_next_tile && pc - nchunk < nchunk) prefetch(tn, pc - nchunk);
      }
      // stage the next chunk in registers
      Unit nxt;
      bool have_nxt = false;
      if (ch + 1 < nchunk) { have_nxt = ok_here; if (have_nxt) load_unit(nxt, fill, sC); fill += 4 * sC; }
      else if (has_next_tile) { have_nxt = ok_next; if (have_nxt) load_unit(nxt, fill_next_tile, sC); }

      const float4* X = buf0 + parity * slots;
#pragma unroll
      for (int k = 0; k < CV_NP; ++k) {
        if (o_warp[k] == nullptr) continue;
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
            const float* q = gsrc + (int64_t)(ch * 4 + c) * sC + o;
            float acc = __fmul_rn(__ldg(q), w_nw[k]);
            acc = fmaf(__ldg(q + dx), w_ne[k], acc);
            acc = fmaf(__ldg(q + dy), w_sw[k], acc);
            acc = fmaf(__ldg(q + dy + dx), w_se[k], acc);
            t[c] = acc;
          }
          r = make_float4(t[0], t[1], t[2], t[3]);
        }
        const float4 own = X[s_own[k]];
        float* ow = o_warp[k];
        float* op = ow + pass_delta;
        ow[0] = r.x; op[0] = own.x; ow += cs; op += cs;
        ow[0] = r.y; op[0] = own.y; ow += cs; op += cs;
        ow[0] = r.z; op[0] = own.z; ow += cs; op += cs;
        ow[0] = r.w; op[0] = own.w;
        o_warp[k] = ow + cs;
      }
      if (have_nxt) store_unit(buf0 + (parity ^ 1) * slots, u_p4, nxt);
      parity ^= 1;
      __syncthreads();
    }
    ti = tn;
  }
}

// ------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------
constexpr int CV_KO = 2;   // extra gather-list entries per target kept in shared memory (beyond CV_K)

__device__ __forceinline__ bool cv_near_row(int sy, int ty, const CvGeom& g) {
  const int h0 = (ty / g.R) * g.R;
  return sy >= h0 - g.HALO && sy < h0 + g.R + g.HALO;
}
__device__ __forceinline__ bool cv_near_col(int dx, const CvGeom& g) { return dx >= -g.DCAP && dx <= g.DCAP; }

// axis_coord with the division by a power-of-two size done as an exact multiplication
__device__ __forceinline__ Axis axis_coord_fast(float base, float flow, int size, float inv, bool pow2) {
  if (!pow2) return axis_coord(base, flow, size);
  float gq = __fadd_rn(base, __fmul_rn(flow, inv));       // flow / 2^k == flow * 2^-k exactly
  const bool pass_clamp = (gq >= -1.f) && (gq <= 1.f);
  gq = fminf(fmaxf(gq, -1.f), 1.f);
  const float hi = (float)(size - 1);
  float i = __fmul_rn(__fmul_rn(__fadd_rn(gq, 1.f), 0.5f), hi);
  const bool pass_clip = (i > 0.f) && (i < hi);
  i = fminf(hi, fmaxf(i, 0.f));
  Axis a;
  a.i = i; a.i0f = floorf(i); a.i0 = (int)a.i0f;
  a.gmult = (pass_clamp && pass_clip) ? 1.f : 0.f;
  return a;
}

__global__ void __launch_bounds__(CV_THREADS, CV_CTAS)
warp_bwd_cvec_kernel(const float* __restrict__ gout, const float* __restrict__ x1, const float* __restrict__ x2,
                     int64_t sB, int64_t sC, const float* __restrict__ flow, const float* __restrict__ xs,
                     const float* __restrict__ ys, float* __restrict__ gx1, float* __restrict__ gx2,
                     float* __restrict__ gflow, int C, int H, int W, CvGeom g) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int slots = g.WR * W;
  const uint32_t buf_bytes = (uint32_t)(2 * slots * sizeof(float4));   // one stage: x window | gout window
  const uint32_t g_off = (uint32_t)(slots * sizeof(float4));
  // [stage 0][stage 1][ix][iy][anchor][overflow weights][overflow offsets][gates][range]
  float* s_ix = reinterpret_cast<float*>(smem_raw + 2 * buf_bytes);
  float* s_iy = s_ix + slots;
  int* s_an = reinterpret_cast<int*>(s_iy + slots);                    // (y0 << 16) | x0 of every window pixel
  float* s_ovw = reinterpret_cast<float*>(s_an + slots);               // [CV_KO][CV_NP][CV_THREADS]
  int* s_ovo = reinterpret_cast<int*>(s_ovw + CV_KO * CV_NP * CV_THREADS);
  int* s_rng = s_ovo + CV_KO * CV_NP * CV_THREADS;
  float* s_pass = reinterpret_cast<float*>(s_rng + 4);                  // [2 stages][4 channels][R*W], cp.async filled
  float* s_flow = s_pass + 2 * 4 * g.R * W;                             // [2 components][WR*W] of the tile being prepared
  unsigned char* s_gate = reinterpret_cast<unsigned char*>(s_flow + 2 * slots);  // core pixels: bit0 = x gate, bit1 = y gate
  const int tid = threadIdx.x;
  const int nchunk = C >> 2;
  const int64_t cs = (int64_t)4 * HW;
  const bool w_pow2 = (W & (W - 1)) == 0, h_pow2 = (H & (H - 1)) == 0;
  const float inv_w = 1.f / (float)W, inv_h = 1.f / (float)H;
  const int u_p4 = 4 * tid;
  const int u_row = u_p4 / W, u_col = u_p4 - u_row * W;
  const bool u_any = tid < g.U;
  int p_r[CV_NP], p_c[CV_NP];
#pragma unroll
  for (int k = 0; k < CV_NP; ++k) {
    const int pl = tid + k * CV_THREADS;
    p_r[k] = pl / W; p_c[k] = pl - p_r[k] * W;
  }

  int tile = blockIdx.x;
  if (tile >= g.ntiles) return;
  TileId ti = tile_of(tile, g);
  auto x_ptr = [&](const TileId& t, int ch) -> const float* {
    return (t.t ? x2 : x1) + t.b * sB + (int64_t)(ch * 4) * sC + (int64_t)(t.h0 - g.HALO + u_row) * W + u_col;
  };
  auto g_ptr = [&](const TileId& t, int ch) -> const float* {
    return gout + ((int64_t)(t.b * C + ch * 4) * 4 + 1 + t.t) * HW + (int64_t)(t.h0 - g.HALO + u_row) * W + u_col;
  };
  auto unit_ok = [&](const TileId& t) { const int y = t.h0 - g.HALO + u_row; return u_any && y >= 0 && y < H; };
  auto prefetch = [&](const TileId& t, int ch) {  // lanes 0..11: x window, gout window, gout pass rows; 12..13: flow
    const int r_lo = max(0, t.h0 - g.HALO), r_hi = min(H, t.h0 + g.R + g.HALO);
    const uint32_t wbytes = (uint32_t)((r_hi - r_lo) * W * sizeof(float));
    const int c = ch * 4 + (tid & 3);
    if (tid < 4) {
      bulk_prefetch_l2((t.t ? x2 : x1) + t.b * sB + (int64_t)c * sC + r_lo * W, wbytes);
    } else if (tid < 8) {
      bulk_prefetch_l2(gout + ((int64_t)(t.b * C + c) * 4 + 1 + t.t) * HW + r_lo * W, wbytes);
    } else if (tid < 12) {
      bulk_prefetch_l2(gout + ((int64_t)(t.b * C + c) * 4 + (t.t ? 3 : 0)) * HW + t.h0 * W,
                       (uint32_t)((min(H, t.h0 + g.R) - t.h0) * W * sizeof(float)));
    } else if (ch == 0 && tid < 14) {
      bulk_prefetch_l2(flow + ((int64_t)(t.b * 2 + (tid - 12)) * 2 + t.t) * HW + r_lo * W, wbytes);
    }
  };
  // cp.async (16 B, no registers): gout pass-slot rows of one chunk -> s_pass[stage]; flow rows of a tile -> s_flow
  const int core = g.R * W;
  auto copy_pass = [&](const TileId& t, int ch, int stage) {
    const int rows = min(g.R, H - t.h0);
    const int upc = rows * W / 4;                                 // 16-byte units per channel
    const float* src = gout + ((int64_t)(t.b * C + ch * 4) * 4 + (t.t ? 3 : 0)) * HW + t.h0 * W;
    float* dst = s_pass + stage * 4 * core;
    for (int c = 0; c < 4; ++c)
      for (int j = tid; j < upc; j += CV_THREADS) cp_async16(dst + c * core + 4 * j, src + c * cs + 4 * j);
  };
  auto copy_flow = [&](const TileId& t) {
    const int r_lo = max(0, t.h0 - g.HALO), r_hi = min(H, t.h0 + g.R + g.HALO);
    const int units = (r_hi - r_lo) * W / 4, woff = (r_lo - (t.h0 - g.HALO)) * W;
    for (int comp = 0; comp < 2; ++comp) {
      const float* src = flow + ((int64_t)(t.b * 2 + comp) * 2 + t.t) * HW + r_lo * W;
      for (int j = tid; j < units; j += CV_THREADS) cp_async16(s_flow + comp * slots + woff + 4 * j, src + 4 * j);
    }
  };
  {
    copy_pass(ti, 0, 0);
    copy_flow(ti);
    cp_async_commit();
    if (tid < 12)
      for (int d = 0; d <= g.PF && d < nchunk; ++d) prefetch(ti, d);
    if (unit_ok(ti)) {
      Unit vx, vg;
      load_unit(vx, x_ptr(ti, 0), sC);
      load_unit(vg, g_ptr(ti, 0), cs);
      store_unit(reinterpret_cast<float4*>(smem_raw), u_p4, vx);
      store_unit(reinterpret_cast<float4*>(smem_raw + g_off), u_p4, vg);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  int pstage = 0;                                // s_pass stage being consumed
  unsigned char* cur = smem_raw;                 // stage being consumed
  unsigned char* oth = smem_raw + buf_bytes;     // stage being filled
  for (; tile < g.ntiles; tile += gridDim.x) {
    const int ntile = tile + gridDim.x;
    const bool has_next_tile = ntile < g.ntiles;
    const TileId tn = has_next_tile ? tile_of(ntile, g) : ti;
    const int wr0 = ti.h0 - g.HALO;
    const int wlo = max(0, wr0), whi = min(H, ti.h0 + g.R + g.HALO);
    const float* fl = flow + ((int64_t)(ti.b * 2) * 2 + ti.t) * HW;
    // ---------------- phase 0: sample coordinates of the window + displacement range ----------------
    if (tid < 4) s_rng[tid] = (tid & 1) ? -(1 << 30) : (1 << 30);
    __syncthreads();
    {
      int mn_x = 1 << 30, mx_x = -(1 << 30), mn_y = 1 << 30, mx_y = -(1 << 30);
      int r = tid / W, col = tid - r * W;
      const int dr = CV_THREADS / W, dc = CV_THREADS - dr * W;
      for (int i = tid; i < slots; i += CV_THREADS) {
        const int sy = wr0 + r;
        if (sy >= wlo && sy < whi) {
          const int p = sy * W + col;
          const Axis ax = axis_coord_fast(__ldg(xs + col), s_flow[i], W, inv_w, w_pow2);
          const Axis ay = axis_coord_fast(__ldg(ys + sy), s_flow[slots + i], H, inv_h, h_pow2);
          s_ix[i] = ax.i; s_iy[i] = ay.i; s_an[i] = (ay.i0 << 16) | ax.i0;
          const int cr = r - g.HALO;
          if (cr >= 0 && cr < g.R) s_gate[cr * W + col] = (unsigned char)((ax.gmult != 0.f ? 1 : 0) | (ay.gmult != 0.f ? 2 : 0));
          const int dx = ax.i0 - col, dy = ay.i0 - sy;
          mn_x = min(mn_x, dx); mx_x = max(mx_x, dx); mn_y = min(mn_y, dy); mx_y = max(mx_y, dy);
        }
        r += dr; col += dc;
        if (col >= W) { col -= W; ++r; }
      }
      mn_x = __reduce_min_sync(0xffffffffu, mn_x); mx_x = __reduce_max_sync(0xffffffffu, mx_x);
      mn_y = __reduce_min_sync(0xffffffffu, mn_y); mx_y = __reduce_max_sync(0xffffffffu, mx_y);
      if ((tid & 31) == 0) {
        atomicMin(s_rng + 0, mn_x); atomicMax(s_rng + 1, mx_x);
        atomicMin(s_rng + 2, mn_y); atomicMax(s_rng + 3, mx_y);
      }
    }
    __syncthreads();
    const int dxlo = max(s_rng[0], -g.DCAP), dxhi = min(s_rng[1], g.DCAP);
    const int dylo = s_rng[2], dyhi = s_rng[3];

    // visit every source of target (ty,tx) in fixed order: f(weight, byte offset of the source slot, hit index)
    auto probe = [&](int ty, int tx, auto&& f) {
      int n = 0;
      const int sy_a = max(wlo, ty - 1 - dyhi), sy_b = min(whi - 1, ty - dylo);
      const int sx_a = max(0, tx - 1 - dxhi), sx_b = min(W - 1, tx - dxlo);
      for (int sy = sy_a; sy <= sy_b; ++sy) {
        const int rowo = (sy - wr0) * W;
        for (int sx = sx_a; sx <= sx_b; ++sx) {
          const int an = s_an[rowo + sx];
          const int x0 = an & 0xffff, y0 = an >> 16;
          const int ex = tx - x0, ey = ty - y0;
          if ((unsigned)ex > 1u || (unsigned)ey > 1u) continue;
          if (!cv_near_col(x0 - sx, g)) continue;
          const float ix = s_ix[rowo + sx], iy = s_iy[rowo + sx];
          const float x0f = (float)x0, y0f = (float)y0;
          const float wx = ex ? __fsub_rn(ix, x0f) : __fsub_rn(__fadd_rn(x0f, 1.f), ix);
          const float wy = ey ? __fsub_rn(iy, y0f) : __fsub_rn(__fadd_rn(y0f, 1.f), iy);
          f(__fmul_rn(wx, wy), swz(rowo + sx) * 16, n);
          ++n;
        }
      }
      return n;
    };

    // ---------------- phase 1: gather lists (targets) and footprints (sources) ----------------
    float lw[CV_NP][CV_K];
    int lo[CV_NP][CV_K];                 // byte offsets of the gather sources inside a window buffer
    int ln[CV_NP], tpix[CV_NP], o_own[CV_NP];
    float wx0[CV_NP], wx1[CV_NP], wy0[CV_NP], wy1[CV_NP];
    float2 gixp[CV_NP], giyp[CV_NP];
    int o_nw[CV_NP], o_ne[CV_NP], o_sw[CV_NP], o_se[CV_NP], gofs[CV_NP];
#pragma unroll
    for (int k = 0; k < CV_NP; ++k) {
      const int h = ti.h0 + p_r[k];
      const int ownp = (p_r[k] + g.HALO) * W + p_c[k];
      gixp[k] = giyp[k] = f2(0.f, 0.f);
      o_own[k] = swz(ownp) * 16;
#pragma unroll
      for (int j = 0; j < CV_K; ++j) { lw[k][j] = 0.f; lo[k][j] = o_own[k]; }
      tpix[k] = -1; ln[k] = 0; o_nw[k] = -1; o_ne[k] = o_sw[k] = o_se[k] = 0; gofs[k] = 0;
      wx0[k] = wx1[k] = wy0[k] = wy1[k] = 0.f;
      if (p_r[k] < g.R && h < H) {
        tpix[k] = h * W + p_c[k];
        ln[k] = probe(h, p_c[k], [&](float w, int off, int n) {
#pragma unroll
          for (int j = 0; j < CV_K; ++j)
            if (n == j) { lw[k][j] = w; lo[k][j] = off; }
          if (n >= CV_K && n < CV_K + CV_KO) {
            s_ovw[((n - CV_K) * CV_NP + k) * CV_THREADS + tid] = w;
            s_ovo[((n - CV_K) * CV_NP + k) * CV_THREADS + tid] = off;
          }
        });
        // own pixel as a source: footprint from the phase-0 tables
        const float ix = s_ix[ownp], iy = s_iy[ownp];
        const int an = s_an[ownp];
        const int x0 = an & 0xffff, y0 = an >> 16;
        const float x0f = (float)x0, y0f = (float)y0;
        wx0[k] = __fsub_rn(__fadd_rn(x0f, 1.f), ix); wx1[k] = __fsub_rn(ix, x0f);
        wy0[k] = __fsub_rn(__fadd_rn(y0f, 1.f), iy); wy1[k] = __fsub_rn(iy, y0f);
        const bool x1ok = x0 + 1 <= W - 1, y1ok = y0 + 1 <= H - 1;
        gofs[k] = (y0 * W + x0) * 4 + (x1ok ? 1 : 0) + (y1ok ? 2 : 0);
        const int sr = y0 - wr0;
        const int dx = x1ok ? 1 : 0, dy = y1ok ? W : 0;
        if (sr >= 0 && sr + (y1ok ? 1 : 0) < g.WR) {
          const int q = sr * W + x0;
          o_nw[k] = swz(q) * 16; o_ne[k] = swz(q + dx) * 16; o_sw[k] = swz(q + dy) * 16; o_se[k] = swz(q + dy + dx) * 16;
        }
      }
    }
    float* gxo = (ti.t ? gx2 : gx1) + ti.b * sB;
    const float* xg = (ti.t ? x2 : x1) + ti.b * sB;
    const bool ok_here = unit_ok(ti), ok_next = unit_ok(tn);
    const float* fill_x = x_ptr(ti, 1);
    const float* fill_g = g_ptr(ti, 1);

    // ---------------- phase 2: chunks of 4 channels ----------------
    for (int ch = 0; ch < nchunk; ++ch) {
      if (tid < 14) {
        const int pc = ch + 1 + g.PF;
        if (pc < nchunk) prefetch(ti, pc);
        else if (has_next_tile && pc - nchunk < nchunk) prefetch(tn, pc - nchunk);
      }
      Unit nx, ng;
      bool have_nxt = false;
      if (ch + 1 < nchunk) {
        have_nxt = ok_here;
        if (have_nxt) { load_unit(nx, fill_x, sC); load_unit(ng, fill_g, cs); }
        fill_x += 4 * sC; fill_g += 4 * cs;
      } else if (has_next_tile) {
        have_nxt = ok_next;
        if (have_nxt) { load_unit(nx, x_ptr(tn, 0), sC); load_unit(ng, g_ptr(tn, 0), cs); }
      }
      if (ch + 1 < nchunk) copy_pass(ti, ch + 1, pstage ^ 1);
      else if (has_next_tile) { copy_pass(tn, 0, pstage ^ 1); copy_flow(tn); }
      cp_async_commit();
      const unsigned char* Xb = cur;
      const unsigned char* Gb = cur + g_off;
      const float* Pb = s_pass + pstage * 4 * core;
#pragma unroll
      for (int k = 0; k < CV_NP; ++k) {
        if (tpix[k] < 0) continue;
        const float* gp = Pb + tid + k * CV_THREADS;
        const float p0 = gp[0], p1 = gp[core], p2 = gp[2 * core], p3 = gp[3 * core];
        // target side: gather the scatter (zero-weight padding instead of predicates)
        float2 a01 = f2(0.f, 0.f), a23 = f2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < CV_K; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(Gb + lo[k][j]);
          a01 = __ffma2_rn(bc(lw[k][j]), f2(v.x, v.y), a01);
          a23 = __ffma2_rn(bc(lw[k][j]), f2(v.z, v.w), a23);
        }
        if (ln[k] > CV_K) {                       // rare: entries CV_K .. CV_K+CV_KO-1 live in shared memory
          const int ne = min(ln[k], CV_K + CV_KO) - CV_K;
          for (int j = 0; j < ne; ++j) {
            const float w = s_ovw[(j * CV_NP + k) * CV_THREADS + tid];
            const float4 v = *reinterpret_cast<const float4*>(Gb + s_ovo[(j * CV_NP + k) * CV_THREADS + tid]);
            a01 = __ffma2_rn(bc(w), f2(v.x, v.y), a01);
            a23 = __ffma2_rn(bc(w), f2(v.z, v.w), a23);
          }
          if (ln[k] > CV_K + CV_KO) {             // very rare: strongly converging flow, re-probe
            probe(ti.h0 + p_r[k], p_c[k], [&](float w, int off, int n) {
              if (n >= CV_K + CV_KO) {
                const float4 v = *reinterpret_cast<const float4*>(Gb + off);
                a01 = __ffma2_rn(bc(w), f2(v.x, v.y), a01);
                a23 = __ffma2_rn(bc(w), f2(v.z, v.w), a23);
              }
            });
          }
        }
        // source side: flow-gradient sums, factored:  gix += go * (wy0 (ne-nw) + wy1 (se-sw)),
        //                                             giy += go * (wx0 (sw-nw) + wx1 (se-ne))
        // (an out-of-bounds tap aliases an in-bounds slot; its weight wx1 / wy1 is exactly 0 there and the
        // other axis is gated off, so the sums equal ATen's skip-the-tap form)
        const float4 go = *reinterpret_cast<const float4*>(Gb + o_own[k]);
        float4 vnw, vne, vsw, vse;
        if (o_nw[k] >= 0) {
          vnw = *reinterpret_cast<const float4*>(Xb + o_nw[k]); vne = *reinterpret_cast<const float4*>(Xb + o_ne[k]);
          vsw = *reinterpret_cast<const float4*>(Xb + o_sw[k]); vse = *reinterpret_cast<const float4*>(Xb + o_se[k]);
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
      gxo += 4 * sC; xg += 4 * sC;
      if (have_nxt) {
        store_unit(reinterpret_cast<float4*>(oth), u_p4, nx);
        store_unit(reinterpret_cast<float4*>(oth + g_off), u_p4, ng);
      }
      unsigned char* tmp = cur; cur = oth; oth = tmp;
      pstage ^= 1;
      cp_async_wait_all();
      __syncthreads();
    }
    // ---------------- phase 3: flow gradient of this tile ----------------
#pragma unroll
    for (int k = 0; k < CV_NP; ++k) {
      if (tpix[k] < 0) continue;
      const int gate = s_gate[p_r[k] * W + p_c[k]];
      const int64_t fo = ((int64_t)(ti.b * 2) * 2 + ti.t) * HW + tpix[k];
      const float mx = __fmul_rn((gate & 1) ? 1.f : 0.f, __fmul_rn((float)(W - 1), 0.5f));
      const float my = __fmul_rn((gate & 2) ? 1.f : 0.f, __fmul_rn((float)(H - 1), 0.5f));
      gflow[fo] = __fdiv_rn(__fmul_rn(mx, __fadd_rn(gixp[k].x, gixp[k].y)), (float)W);
      gflow[fo + 2 * (int64_t)HW] = __fdiv_rn(__fmul_rn(my, __fadd_rn(giyp[k].x, giyp[k].y)), (float)H);
    }
    ti = tn;
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
  int R = (CV_THREADS * CV_NP) / W;
  if (R < 1) return false;
  if (R > H) R = H;
  g.R = R;
  g.HALO = halo < 1 ? 1 : halo;
  while (g.HALO > 1 && (R + 2 * g.HALO) * W / 4 > CV_THREADS) --g.HALO;   // one fill unit per thread per array
  g.DCAP = g.HALO + 1;
  g.WR = R + 2 * g.HALO;
  if ((g.WR * W) % 8 != 0) return false;
  g.U = g.WR * W / 4;
  if (g.U > CV_THREADS) return false;
  g.nbands = (H + R - 1) / R;
  g.ntiles = 2 * B * g.nbands;
  g.PF = option(OPT_CVEC_PF);
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
  const int sms = device_info().sms * CV_CTAS;
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
  const size_t smem = 4 * (size_t)g.WR * W * sizeof(float4) + 5 * (size_t)g.WR * W * sizeof(float) +
                      8 * (size_t)g.R * W * sizeof(float) +
                      2 * (size_t)CV_KO * CV_NP * CV_THREADS * sizeof(float) + 4 * sizeof(int) + (size_t)g.R * W;
  static thread_local size_t configured = 0;
  if (int e = cv_prepare(warp_bwd_cvec_kernel, smem, configured)) return e;
  const int sms = device_info().sms * CV_CTAS;
  const int grid = g.ntiles < sms ? g.ntiles : sms;
  warp_bwd_cvec_kernel<<<grid, CV_THREADS, smem, st>>>(gout, x1, x2, sB, sC, flow, xs, ys, gx1, gx2, gflow, C, H, W, g);
  warp_bwd_cvec_far_kernel<<<dim3((H * W + 255) / 256, 2 * B), 256, 0, st>>>(gout, flow, xs, ys, gx1, gx2, sB, sC, C,
                                                                            H, W, g);
  count_launch(2);
  return check_launch("warp_bwd_cvec");
}

}  // namespace smow
