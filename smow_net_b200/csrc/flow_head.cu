// Row N1: the OFW flow head (reference models/SMOW_Net.py:602,606-608 = models/SMOW_Net_LW.py:444,448-450)
//     seg_up = F.interpolate(seg_down, size=(2,H,W), 'trilinear', align_corners=True)     (per-frame bilinear, x8)
//     flow   = flow_make(torch.cat([x, seg_up], 1))                                          Conv3d(2C -> 2, 3x3x3, pad 1, no bias)
// without the up-sampled tensor and without the concat.  The convolution is linear in its two channel halves:
//   * x half: a direct 3x3x3 stencil over the NDHWC feature stack (C channels, both input frames feed both output frames
//     through different temporal taps; the third temporal tap always falls on the zero padding);
//   * seg half: up-sampling and the channel contraction commute, so the contraction runs at LOW resolution,
//         Z[b, tap, i, j, (t,o)] = sum_{t', c} W[o, C + c, t' - t + 1, kh, kw] * seg_down[b, c, t', i, j]        (host: one tiny einsum)
//     and the kernel adds  sum_tap [p + tap inside the image] * bilerp(Z[b, tap], p + tap)  — 9 bilinear look-ups into a
//     36 KB table per pair that lives in shared memory, instead of streaming a (B, C, 2, H, W) tensor.
// HBM traffic of the forward: x once (2*C*HW*4 bytes per pair) + the flow (16*HW) — the reference moves ~3x the size of
// x more (write seg_up, read x + seg_up, write the 2C-channel concat, read it again).
// Backward: d x (transposed stencil, gather form), d Z (transposed bilinear, separable, deterministic) and d W_x
// (per-CTA partial sums + ordered reduction); d seg_down and d W_seg follow from d Z through the einsum's autograd.
// fp32, x NDHWC (channels_last_3d), flow contiguous (B,2,2,H,W) = [b][o][t][y][x].  Bandwidth / FP32-issue bound, N = 2
// output channels: nothing here for tensor cores.
#include "common.cuh"
#include "bulk.cuh"

namespace smow {

// 4-byte asynchronous copy global -> shared with zero fill when !valid (all copies of a tile are in flight at once)
__device__ __forceinline__ void cp_async4_zfill(float* smem_dst, const float* gmem_src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(valid ? 4 : 0) : "memory");
}
// gradient tile with a one-row / four-column zero halo: gs[q][r][4 + x] = g[b][o][t][y0 - 1 + r][x]   (q = 2t + o)
__device__ __forceinline__ void load_grad_tile(float* gs, const float* __restrict__ g, int b, int y0, int R2, int LD, int H, int W) {
  const int n = 4 * R2 * LD;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int col = i % LD - 4, rr = i / LD, r = rr % R2, q = rr / R2;
    const int yy = y0 + r - 1, t = q >> 1, o = q & 1;
    const bool ok = col >= 0 && col < W && yy >= 0 && yy < H;
    cp_async4_zfill(gs + i, ok ? g + ((((size_t)b * 2 + o) * 2 + t) * H + yy) * W + col : g, ok);
  }
}

// ATen's align_corners=True source index (UpSample.h area_pixel_compute_source_index): src = dst * (in-1)/(out-1)
struct UpIdx { int i0, i1; float l0, l1; };
__device__ __forceinline__ UpIdx up_index(int dst, int in_size, float scale) {
  const float r = scale * (float)dst;
  UpIdx u;
  u.i0 = (int)r;
  if (u.i0 > in_size - 1) u.i0 = in_size - 1;
  u.i1 = u.i0 + ((u.i0 < in_size - 1) ? 1 : 0);
  u.l1 = r - (float)u.i0;
  u.l0 = 1.f - u.l1;
  return u;
}
__host__ __device__ inline float up_scale(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}

// weights of the x half re-ordered for the kernels: wx[t'][tap][t][o][c]  (tap = kh*3 + kw; kt = t' - t + 1)
__global__ void flow_head_pack_kernel(const float* __restrict__ w, float* __restrict__ wx, int C) {
  const int n = 2 * 9 * 4 * C;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int c = e % C, to = (e / C) % 4, tap = (e / (4 * C)) % 9, tp = e / (36 * C);
    const int t = to >> 1, o = to & 1, kt = tp - t + 1;
    wx[e] = w[(((size_t)o * 2 * C + c) * 3 + kt) * 9 + tap];            // w: (2, 2C, 3, 3, 3)
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------------
// thread = 4 consecutive pixels of one row (all four (t, o) outputs); CTA = rows_per_cta rows of one pair.
template <int C>
__global__ void __launch_bounds__(256)
flow_head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wxp, const float* __restrict__ z,
                     float* __restrict__ flow, int H, int W, int h, int w, int rows_per_cta) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                                   // [2][9][4][C]
  float* zs = ws + 72 * C;                          // [9][h*w][4]
  const int b = blockIdx.y, y0 = blockIdx.x * rows_per_cta;
  for (int i = threadIdx.x; i < 72 * C / 4; i += 256) cp_async16(reinterpret_cast<float4*>(ws) + i, reinterpret_cast<const float4*>(wxp) + i);
  const float4* zb = reinterpret_cast<const float4*>(z) + (size_t)b * 9 * h * w;
  for (int i = threadIdx.x; i < 9 * h * w; i += 256) cp_async16(reinterpret_cast<float4*>(zs) + i, zb + i);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  const int quads = W >> 2;
  const float sy = up_scale(h, H), sx = up_scale(w, W);
  for (int item = threadIdx.x; item < rows_per_cta * quads; item += 256) {
    const int y = y0 + item / quads, xq = (item % quads) * 4;
    if (y >= H) continue;
    float acc[4][4];                                // [pixel][t*2+o]
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
    // ---- x half: rows y-1..y+1, columns xq-1..xq+4 of both frames; iterations (t', kh, channel vector) are flattened and the
    // six taps of the NEXT iteration are loaded while the current ones are consumed
    constexpr int Q4 = C / 4, NIT = 6 * Q4;
    auto load6 = [&](int it, float4 (&xv)[6]) {
      const int cv = it % Q4, kh = (it / Q4) % 3, tp = it / (3 * Q4), yy = y + kh - 1;
      const bool rowok = yy >= 0 && yy < H;
      const float* row = x + (((size_t)(b * 2 + tp) * H + (rowok ? yy : 0)) * W) * C;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int xx = xq + j - 1;
        xv[j] = (rowok && xx >= 0 && xx < W) ? __ldg(reinterpret_cast<const float4*>(row + (size_t)xx * C) + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 cur[6], nxt[6];
    load6(0, cur);
    for (int it = 0; it < NIT; ++it) {
      if (it + 1 < NIT) load6(it + 1, nxt);
      const int cv = it % Q4, kh = (it / Q4) % 3, tp = it / (3 * Q4);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* wv = reinterpret_cast<const float4*>(ws + ((tp * 9 + kh * 3 + kw) * 4) * C) + cv;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wq = wv[q * Q4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float4 v = cur[p + kw];
            acc[p][q] = fmaf(v.x, wq.x, fmaf(v.y, wq.y, fmaf(v.z, wq.z, fmaf(v.w, wq.w, acc[p][q]))));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) cur[j] = nxt[j];
    }
    // ---- seg half: 9 bilinear look-ups into the low-resolution table
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + kh - 1;
      if (yy < 0 || yy >= H) continue;
      const UpIdx uy = up_index(yy, h, sy);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int xx = xq + j - 1;
        if (xx < 0 || xx >= W) continue;
        const UpIdx ux = up_index(xx, w, sx);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int p = j - kw;                      // pixel of the quad that sees column xx through tap kw
          if (p < 0 || p > 3) continue;
          const float4* zt = reinterpret_cast<const float4*>(zs) + (size_t)(kh * 3 + kw) * h * w;
          const float4 a = zt[uy.i0 * w + ux.i0], bq = zt[uy.i0 * w + ux.i1], c = zt[uy.i1 * w + ux.i0], d = zt[uy.i1 * w + ux.i1];
          const float w00 = uy.l0 * ux.l0, w01 = uy.l0 * ux.l1, w10 = uy.l1 * ux.l0, w11 = uy.l1 * ux.l1;
          acc[p][0] += w00 * a.x + w01 * bq.x + w10 * c.x + w11 * d.x;
          acc[p][1] += w00 * a.y + w01 * bq.y + w10 * c.y + w11 * d.y;
          acc[p][2] += w00 * a.z + w01 * bq.z + w10 * c.z + w11 * d.z;
          acc[p][3] += w00 * a.w + w01 * bq.w + w10 * c.w + w11 * d.w;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {                   // q = t*2 + o  ->  flow[b][o][t][y][x]
      const int t = q >> 1, o = q & 1;
      float* dst = flow + ((((size_t)b * 2 + o) * 2 + t) * H + y) * W + xq;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
    }
  }
}

// ---- backward: d x ------------------------------------------------------------------------------------------------------
// gx[b,t',y,x,c] = sum_{t,o,kh,kw} W[o,c,t'-t+1,kh,kw] * g[b,o,t,y-kh+1,x-kw+1].  thread = (4 consecutive pixels, 4 channels, both t').
template <int C>
__global__ void __launch_bounds__(256)
flow_head_bwd_x_kernel(const float* __restrict__ g, const float* __restrict__ wxp, float* __restrict__ gx, int H, int W,
                       int rows_per_cta) {
  extern __shared__ __align__(16) float sm[];
  float* ws = sm;                                   // [2][9][4][C]
  float* gs = ws + 72 * C;                          // [4 (t*2+o)][rows_per_cta + 2][W + 8]   (column x stored at x + 4)
  const int b = blockIdx.y, y0 = blockIdx.x * rows_per_cta, LD = W + 8, R2 = rows_per_cta + 2;
  for (int i = threadIdx.x; i < 72 * C / 4; i += 256) cp_async16(reinterpret_cast<float4*>(ws) + i, reinterpret_cast<const float4*>(wxp) + i);
  load_grad_tile(gs, g, b, y0, R2, LD, H, W);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  constexpr int Q = C / 4;
  const int quads = W >> 2;
  for (int item = threadIdx.x; item < rows_per_cta * quads * Q; item += 256) {
    const int cv = item % Q, qd = (item / Q) % quads, r = item / (Q * quads);
    const int y = y0 + r, xq = qd * 4;
    if (y >= H) continue;
    float4 acc[2][4];
#pragma unroll
    for (int tp = 0; tp < 2; ++tp)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[tp][p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // source row y - kh + 1 is tile row r + 1 - kh + 1 ... stored with a one-row halo: index (r + 2 - kh)
        const float* grow = gs + (q * R2 + (r + 2 - kh)) * LD + 4 + xq;
        float gv[6];                                 // columns xq-1 .. xq+4
#pragma unroll
        for (int j = 0; j < 6; ++j) gv[j] = grow[j - 1];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int tp = 0; tp < 2; ++tp) {
            const float4 wq = reinterpret_cast<const float4*>(ws + (((tp * 9 + kh * 3 + kw) * 4) + q) * C)[cv];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float gg = gv[p - kw + 2];       // g at column (xq + p) - kw + 1
              acc[tp][p].x = fmaf(gg, wq.x, acc[tp][p].x); acc[tp][p].y = fmaf(gg, wq.y, acc[tp][p].y);
              acc[tp][p].z = fmaf(gg, wq.z, acc[tp][p].z); acc[tp][p].w = fmaf(gg, wq.w, acc[tp][p].w);
            }
          }
        }
      }
    }
#pragma unroll
    for (int tp = 0; tp < 2; ++tp)
#pragma unroll
      for (int p = 0; p < 4; ++p)
        reinterpret_cast<float4*>(gx + (((size_t)(b * 2 + tp) * H + y) * W + xq + p) * C)[cv] = acc[tp][p];
  }
}

// ---- backward: d Z (transposed bilinear of the nine shifted gradient planes) ------------------------------------------
// gz[b,tap,i,j,q] = sum_{P,Q} Ry[P][i] Rx[Q][j] [P-kh+1, Q-kw+1 inside] g[b,q][P-kh+1][Q-kw+1]     (P,Q = sampling position)
// one CTA per (b, q, kh): U[i][Q'] = sum_P Ry[P][i] g[P-kh+1][Q'] (Q' = source column), then the column pass for kw = 0..2.
// The bilinear source index / weights of every sampling row and column are tabulated once in shared memory.
__global__ void __launch_bounds__(256)
flow_head_bwd_z_kernel(const float* __restrict__ g, float* __restrict__ gz, int H, int W, int h, int w) {
  extern __shared__ __align__(16) float sm[];
  float* gp = sm;                                   // [H][W] gradient plane
  float* U = gp + H * W;                            // [h][W]
  float* l1y = U + h * W;                           // [H] weight of the second source row
  float* l1x = l1y + H;                             // [W]
  int* i0y = reinterpret_cast<int*>(l1x + W);       // [H] first source row | second << 16
  int* i0x = i0y + H;                               // [W]
  const int kh = blockIdx.x % 3, q = (blockIdx.x / 3) & 3, b = blockIdx.x / 12, t = q >> 1, o = q & 1;
  const float* src = g + (((size_t)b * 2 + o) * 2 + t) * H * W;
  for (int i = threadIdx.x; i < H * W / 4; i += 256) cp_async16(reinterpret_cast<float4*>(gp) + i, reinterpret_cast<const float4*>(src) + i);
  cp_async_commit();
  const float sy = up_scale(h, H), sx = up_scale(w, W);
  for (int P = threadIdx.x; P < H; P += 256) { const UpIdx u = up_index(P, h, sy); l1y[P] = u.l1; i0y[P] = u.i0 | (u.i1 << 16); }
  for (int Q = threadIdx.x; Q < W; Q += 256) { const UpIdx u = up_index(Q, w, sx); l1x[Q] = u.l1; i0x[Q] = u.i0 | (u.i1 << 16); }
  cp_async_wait_all();
  __syncthreads();
  for (int e = threadIdx.x; e < h * W; e += 256) {
    const int c = e % W, i = e / W;
    // rows P whose bilinear footprint contains node i lie in [ (i-1)/sy , (i+1)/sy ]
    int lo = sy > 0.f ? (int)floorf((float)(i - 1) / sy) - 1 : 0, hi = sy > 0.f ? (int)ceilf((float)(i + 1) / sy) + 1 : H - 1;
    if (lo < 0) lo = 0;
    if (hi > H - 1) hi = H - 1;
    float acc = 0.f;
    for (int P = lo; P <= hi; ++P) {
      const int ys = P - kh + 1;
      if (ys < 0 || ys >= H) continue;
      const int pk = i0y[P];
      const float l1 = l1y[P];
      float wgt = 0.f;
      if ((pk & 0xffff) == i) wgt += 1.f - l1;
      if ((pk >> 16) == i) wgt += l1;
      if (wgt != 0.f) acc = fmaf(wgt, gp[ys * W + c], acc);
    }
    U[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 3 * h * w; e += 256) {
    const int j = e % w, i = (e / w) % h, kw = e / (w * h);
    int lo = sx > 0.f ? (int)floorf((float)(j - 1) / sx) - 1 : 0, hi = sx > 0.f ? (int)ceilf((float)(j + 1) / sx) + 1 : W - 1;
    if (lo < 0) lo = 0;
    if (hi > W - 1) hi = W - 1;
    float acc = 0.f;
    for (int Q = lo; Q <= hi; ++Q) {
      const int xs = Q - kw + 1;
      if (xs < 0 || xs >= W) continue;
      const int pk = i0x[Q];
      const float l1 = l1x[Q];
      float wgt = 0.f;
      if ((pk & 0xffff) == j) wgt += 1.f - l1;
      if ((pk >> 16) == j) wgt += l1;
      if (wgt != 0.f) acc = fmaf(wgt, U[i * W + xs], acc);
    }
    gz[(((size_t)b * 9 + kh * 3 + kw) * h * w + (size_t)i * w + j) * 4 + q] = acc;
  }
}

// ---- backward: d W_x ---------------------------------------------------------------------------------------------------
// gw[o,c,t'-t+1,kh,kw] = sum_{b,y',x'} g[b,o,t,y'-kh+1,x'-kw+1] * x[b,c,t',y',x']: a [36 combos x pixels] . [pixels x C] product per
// (pair, t', row band).  thread = (6 combos, 4 channels) register tile over every NS-th column of the band; the NS column
// slices are added in a fixed order through shared memory.  part[cta][36][C].
template <int C>
__global__ void __launch_bounds__(256)
flow_head_bwd_w_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ part, int H, int W,
                       int rows_per_cta) {
  extern __shared__ __align__(16) float sm[];
  constexpr int Q = C / 4, TILES = 6 * Q, NS = 256 / TILES;           // C = 16: 24 tiles x 10 slices; 32: 48 x 5; 64: 96 x 2
  const int LD = W + 8, R2 = rows_per_cta + 2;
  float* gs = sm;                                   // [4][R2][LD]
  float* xs = gs + 4 * R2 * LD;                     // [rows][W][C]
  const int bands = (H + rows_per_cta - 1) / rows_per_cta;
  const int band = blockIdx.x % bands, tp = (blockIdx.x / bands) & 1, b = blockIdx.x / (2 * bands);
  const int y0 = band * rows_per_cta;
  const int nrows = H - y0 < rows_per_cta ? H - y0 : rows_per_cta;
  load_grad_tile(gs, g, b, y0, R2, LD, H, W);
  const float4* xsrc = reinterpret_cast<const float4*>(x + (((size_t)(b * 2 + tp) * H + y0) * W) * C);
  for (int i = threadIdx.x; i < nrows * W * C / 4; i += 256) cp_async16(reinterpret_cast<float4*>(xs) + i, xsrc + i);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  const int tile = threadIdx.x % TILES, slice = threadIdx.x / TILES;
  const int cv = tile % Q, cg = tile / Q;                             // combos 6*cg .. 6*cg+5
  float4 acc[6];
  int off[6];                                        // gradient sample of combo a relative to (row r, column xx)
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int combo = 6 * cg + a, tap = combo >> 2, q = combo & 3, kh = tap / 3, kw = tap - kh * 3;
    off[a] = (q * R2 + (2 - kh)) * LD + 5 - kw;
  }
  if (slice < NS) {
    for (int r = 0; r < nrows; ++r) {
      const float* xrow = xs + (size_t)r * W * C + 4 * cv;
      const float* grow = gs + r * LD;
#pragma unroll 2
      for (int xx = slice; xx < W; xx += NS) {
        const float4 xv = *reinterpret_cast<const float4*>(xrow + (size_t)xx * C);
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          const float gg = grow[off[a] + xx];
          acc[a].x = fmaf(gg, xv.x, acc[a].x); acc[a].y = fmaf(gg, xv.y, acc[a].y);
          acc[a].z = fmaf(gg, xv.z, acc[a].z); acc[a].w = fmaf(gg, xv.w, acc[a].w);
        }
      }
    }
  }
  __syncthreads();
  float* red = sm;                                   // [NS][36][C]  (<= 10*36*16*4 = 23 KB)
  if (slice < NS) {
#pragma unroll
    for (int a = 0; a < 6; ++a) reinterpret_cast<float4*>(red + ((size_t)slice * 36 + 6 * cg + a) * C)[cv] = acc[a];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 36 * C; e += 256) {
    float tsum = 0.f;
    for (int k = 0; k < NS; ++k) tsum += red[(size_t)k * 36 * C + e];
    part[(size_t)blockIdx.x * 36 * C + e] = tsum;
  }
}

// gw (2, 2C, 3,3,3): only the x half [:, :C] is written here.  One WARP per (o, c, kt, tap); the temporal tap kt = t' - t + 1
// collects (t',t) = (0,1) for kt = 0, (0,0) and (1,1) for kt = 1, (1,0) for kt = 2.  Lanes stride over the per-CTA partials,
// then a fixed shuffle tree: deterministic.
__global__ void __launch_bounds__(256)
flow_head_bwd_w_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw, int C, int nctas, int bands) {
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= 2 * C * 27) return;
  const int tap = e % 9, kt = (e / 9) % 3, c = (e / 27) % C, o = e / (27 * C);
  const int npairs = nctas / (2 * bands);                              // = B
  float total = 0.f;
  for (int tp = 0; tp < 2; ++tp) {
    const int t = tp - kt + 1;
    if (t < 0 || t > 1) continue;
    const int combo = tap * 4 + t * 2 + o;
    float acc = 0.f;
    for (int k = lane; k < npairs * bands; k += 32) {                  // CTAs are ordered (b, t', band)
      const int bb = k / bands, band = k - bb * bands;
      acc += part[((size_t)((bb * 2 + tp) * bands + band) * 36 + combo) * C + c];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    total += acc;
  }
  if (lane == 0) gw[(((size_t)o * 2 * C + c) * 3 + kt) * 9 + tap] = total;
}

template <typename K> static int opt_in_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return check_launch(what);
  return 0;
}
static bool flow_head_ok(int C, int H, int W, int h, int w) {
  return (C == 16 || C == 32 || C == 64) && W % 4 == 0 && W >= 4 && H >= 1 && h >= 1 && w >= 1 && (size_t)9 * h * w * 16 <= 96 * 1024 &&
         (size_t)H * W * 4 + (size_t)h * W * 4 + 8 * (size_t)(H + W) <= 200 * 1024 && h < 32768 && w < 32768;
}
static int fh_rows_w(int C, int W) {                  // rows per CTA of the d W kernel: x band + gradient halo must fit
  int rows = 8;
  while (rows > 1 && (size_t)rows * W * C * 4 + (size_t)4 * (rows + 2) * (W + 8) * 4 > 160 * 1024) rows >>= 1;
  return rows;
}

}  // namespace smow

using namespace smow;

extern "C" {

int smow_flow_head_supported(int C, int H, int W, int h, int w) { return flow_head_ok(C, H, W, h, w) ? 1 : 0; }

// scratch for the re-ordered x-half weights (72*C floats) + the d W partials
int64_t smow_flow_head_workspace_bytes(int B, int C, int H, int W) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const int rows = fh_rows_w(C, W);
  const int64_t bands = (H + rows - 1) / rows;
  return (int64_t)72 * C * 4 + (int64_t)B * 2 * bands * 36 * C * 4;
}

int smow_flow_head_fwd(const float* x, const float* weight, const float* z, float* flow, int B, int C, int H, int W, int h,
                       int w, void* ws, int64_t ws_bytes, void* stream) {
  if (!x || !weight || !z || !flow || B <= 0) return fail(SMOW_EINVAL, "flow_head: bad shape / null pointer");
  if (!flow_head_ok(C, H, W, h, w)) return fail(SMOW_EDTYPE, "flow_head: unsupported shape C=%d %dx%d coarse %dx%d", C, H, W, h, w);
  if (!aligned16(x) || !aligned16(z) || !aligned16(flow) || !ws || !aligned16(ws) || ws_bytes < (int64_t)72 * C * 4)
    return fail(SMOW_EALIGN, "flow_head: 16 B alignment / workspace of smow_flow_head_workspace_bytes() bytes");
  if (B > 65535) return fail(SMOW_ERANGE, "flow_head: batch too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  float* wxp = reinterpret_cast<float*>(ws);
  flow_head_pack_kernel<<<(72 * C + 255) / 256, 256, 0, st>>>(weight, wxp, C);
  const int rows = 8;
  const size_t smem = (size_t)(72 * C + 9 * h * w * 4) * 4;
  const dim3 grid((H + rows - 1) / rows, B);
  int rc = 0;
  switch (C) {
    case 16: if ((rc = opt_in_smem(flow_head_fwd_kernel<16>, smem, "flow_head_fwd"))) return rc;
             flow_head_fwd_kernel<16><<<grid, 256, smem, st>>>(x, wxp, z, flow, H, W, h, w, rows); break;
    case 32: if ((rc = opt_in_smem(flow_head_fwd_kernel<32>, smem, "flow_head_fwd"))) return rc;
             flow_head_fwd_kernel<32><<<grid, 256, smem, st>>>(x, wxp, z, flow, H, W, h, w, rows); break;
    default: if ((rc = opt_in_smem(flow_head_fwd_kernel<64>, smem, "flow_head_fwd"))) return rc;
             flow_head_fwd_kernel<64><<<grid, 256, smem, st>>>(x, wxp, z, flow, H, W, h, w, rows); break;
  }
  count_launch(2);
  return check_launch("flow_head_fwd");
}

int smow_flow_head_bwd(const float* gflow, const float* x, const float* weight, float* gx, float* gweight, float* gz, int B,
                       int C, int H, int W, int h, int w, void* ws, int64_t ws_bytes, void* stream) {
  if (!gflow || !x || !weight || !gx || !gweight || !gz || B <= 0) return fail(SMOW_EINVAL, "flow_head: bad shape / null pointer");
  if (!flow_head_ok(C, H, W, h, w)) return fail(SMOW_EDTYPE, "flow_head: unsupported shape C=%d %dx%d coarse %dx%d", C, H, W, h, w);
  if (!aligned16(gflow) || !aligned16(x) || !aligned16(gx) || !aligned16(gz) || !ws || !aligned16(ws) ||
      ws_bytes < smow_flow_head_workspace_bytes(B, C, H, W))
    return fail(SMOW_EALIGN, "flow_head: 16 B alignment / workspace of smow_flow_head_workspace_bytes() bytes");
  if (B > 16383) return fail(SMOW_ERANGE, "flow_head: batch too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  float* wxp = reinterpret_cast<float*>(ws);
  float* part = wxp + 72 * C;
  flow_head_pack_kernel<<<(72 * C + 255) / 256, 256, 0, st>>>(weight, wxp, C);
  int rc = 0;
  {   // d x
    const int rows = 8;
    const size_t smem = (size_t)(72 * C + 4 * (rows + 2) * (W + 8)) * 4;
    const dim3 grid((H + rows - 1) / rows, B);
    switch (C) {
      case 16: if ((rc = opt_in_smem(flow_head_bwd_x_kernel<16>, smem, "flow_head_bwd_x"))) return rc;
               flow_head_bwd_x_kernel<16><<<grid, 256, smem, st>>>(gflow, wxp, gx, H, W, rows); break;
      case 32: if ((rc = opt_in_smem(flow_head_bwd_x_kernel<32>, smem, "flow_head_bwd_x"))) return rc;
               flow_head_bwd_x_kernel<32><<<grid, 256, smem, st>>>(gflow, wxp, gx, H, W, rows); break;
      default: if ((rc = opt_in_smem(flow_head_bwd_x_kernel<64>, smem, "flow_head_bwd_x"))) return rc;
               flow_head_bwd_x_kernel<64><<<grid, 256, smem, st>>>(gflow, wxp, gx, H, W, rows); break;
    }
  }
  {   // d Z
    const size_t smem = ((size_t)H * W + (size_t)h * W + 2 * (size_t)(H + W)) * 4;
    if ((rc = opt_in_smem(flow_head_bwd_z_kernel, smem, "flow_head_bwd_z"))) return rc;
    flow_head_bwd_z_kernel<<<B * 12, 256, smem, st>>>(gflow, gz, H, W, h, w);
  }
  {   // d W_x
    const int rows = fh_rows_w(C, W), bands = (H + rows - 1) / rows, nctas = B * 2 * bands;
    size_t smem = ((size_t)4 * (rows + 2) * (W + 8) + (size_t)rows * W * C) * 4;
    const size_t red = (size_t)(256 / (6 * (C / 4))) * 36 * C * 4;   // the slice-reduction scratch re-uses the tile space
    if (smem < red) smem = red;
    switch (C) {
      case 16: if ((rc = opt_in_smem(flow_head_bwd_w_kernel<16>, smem, "flow_head_bwd_w"))) return rc;
               flow_head_bwd_w_kernel<16><<<nctas, 256, smem, st>>>(gflow, x, part, H, W, rows); break;
      case 32: if ((rc = opt_in_smem(flow_head_bwd_w_kernel<32>, smem, "flow_head_bwd_w"))) return rc;
               flow_head_bwd_w_kernel<32><<<nctas, 256, smem, st>>>(gflow, x, part, H, W, rows); break;
      default: if ((rc = opt_in_smem(flow_head_bwd_w_kernel<64>, smem, "flow_head_bwd_w"))) return rc;
               flow_head_bwd_w_kernel<64><<<nctas, 256, smem, st>>>(gflow, x, part, H, W, rows); break;
    }
    flow_head_bwd_w_reduce_kernel<<<(2 * 27 * C + 7) / 8, 256, 0, st>>>(part, gweight, C, nctas, bands);
  }
  count_launch(5);
  return check_launch("flow_head_bwd");
}

}  // extern "C"
