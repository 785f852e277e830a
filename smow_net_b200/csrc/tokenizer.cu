// Row N2: the semantic tokenizer, sole consumer of the warped stack (reference models/SMOW_Net.py:171-190,
// models/SMOW_Net_LW.py:190-209).  Per pair b and frame k of the (B,C,4,H,W) stack:
//     logit[l,p]  = sum_c Wa[l,c] * x[p,c] + ba[l]          (1x1 conv C -> L = 8)
//     attn[l,:]   = softmax over the H*W pixels
//     tokens[l,c] = sum_p attn[l,p] * x[p,c]                (einsum 'bln,bcn->blc')
// The reference runs this as 4 x (strided-frame copy, cuDNN conv, ATen softmax, a K = 16384 / N = C batched GEMM that
// cuBLAS executes at ~1 % of the HBM roofline) plus the matching backward chain.  Here the forward is ONE pass over the
// NDHWC stack (chunk-local softmax statistics + un-normalised partial tokens, combined by a tiny second kernel) and the
// backward is ONE pass that recomputes the attention from the saved (max, sum) and produces d(stack), dWa and dba.
//
// Thread = (pixel, token pair tp, channel group cg): the 8 tokens are split over 4 lanes (2 tokens each), the channels
// over G = C / CG groups of CG = min(C, 16) channels; the LP = 4*G lanes of a pixel are adjacent (lane = cg*4 + tp).
// A lane keeps the 2 x CG weights of ITS tokens and channels in registers for the whole kernel, evaluates only its own
// two exponentials and accumulates a 2 x CG tile of the token matrix — no cross-lane traffic at all for C <= 16, one
// G-lane butterfly per logit beyond.  The work is ~300 FP32 instructions per 64-byte pixel: these kernels are bound by
// FP32 issue, not by HBM.  Everything is summed in a fixed order (no float atomics): results are bit-reproducible.
#include "bulk.cuh"
#include "tokenizer.cuh"

namespace smow {

constexpr int TOK_FWD_THREADS = 256, TOK_BWD_THREADS = 128;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// sum over the G channel-group lanes of a pixel (lane stride 4)
__device__ __forceinline__ float groups_sum(float v, int LP) {
  for (int d = 4; d < LP; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// sum / max over the lanes of a warp that play the same role for different pixels (lane stride LP)
__device__ __forceinline__ float pixels_sum(float v, int LP) {
  for (int d = LP; d < 32; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float pixels_max(float v, int LP) {
  for (int d = LP; d < 32; d <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

template <int CG>
__device__ __forceinline__ void load_row(float (&dst)[CG], const float* p) {
  if constexpr (CG >= 4) {
#pragma unroll
    for (int i = 0; i < CG / 4; ++i) {
      const float4 v = ldg4(p + 4 * i);
      dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
    }
  }
}
template <int CG>
__device__ __forceinline__ void lds_row(float (&dst)[CG], const float* p) {
#pragma unroll
  for (int i = 0; i < CG / 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}
// a staged row as packed pairs, fetched with 16-byte shared loads (half the requests of 8-byte ones)
template <int CG>
__device__ __forceinline__ void lds_pairs(float2 (&dst)[CG / 2], const float* p) {
  if constexpr (CG >= 4) {
#pragma unroll
    for (int i = 0; i < CG / 4; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
      dst[2 * i] = make_float2(v.x, v.y); dst[2 * i + 1] = make_float2(v.z, v.w);
    }
  }
}
// one thread starts the bulk copy of the chunk's rows into shared memory; everybody waits on the mbarrier later
__device__ __forceinline__ void stage_chunk(float* xs, const float* src, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, bytes);
    bulk_g2s(xs, src, bytes, bar);
  }
  __syncthreads();                                       // the barrier's initialisation is visible to every waiter
}
template <int CG>
__device__ __forceinline__ float dot_row(const float (&a)[CG], const float (&b)[CG]) {
  float s0 = 0.f, s1 = 0.f;                              // two chains: instruction-level parallelism
#pragma unroll
  for (int i = 0; i < CG; i += 2) { s0 = fmaf(a[i], b[i], s0); s1 = fmaf(a[i + 1], b[i + 1], s1); }
  return s0 + s1;
}

// block-level, fixed-order reduction of a [2][CG] register tile + 2 scalars per lane role into
// out_mat[l][c] (l = 2*tp + t, c = cg*CG + i) and out_vec[l]
template <int CG, int NT>
__device__ __forceinline__ void reduce_tiles(float (&tile)[2][CG], float (&sc)[2], float* smem, float* out_mat,
                                             float* out_vec, int C, int LP) {
  constexpr int NW = NT / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* smat = smem;                                    // [NW warps][LP][2][CG]
  float* svec = smem + NW * LP * 2 * CG;                 // [NW warps][LP][2]
#pragma unroll
  for (int t = 0; t < 2; ++t) {
#pragma unroll
    for (int i = 0; i < CG; ++i) {
      const float v = pixels_sum(tile[t][i], LP);
      if (lane < LP) smat[((warp * LP + lane) * 2 + t) * CG + i] = v;
    }
    const float v = pixels_sum(sc[t], LP);
    if (lane < LP) svec[(warp * LP + lane) * 2 + t] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TOK_L * C; i += NT) {    // i = l*C + c
    const int l = i / C, c = i - l * C;
    const int li = (c / CG) * 4 + (l >> 1);              // lane role that owns (token l, channel c)
    float t = 0.f;
    for (int wp = 0; wp < NW; ++wp) t += smat[((wp * LP + li) * 2 + (l & 1)) * CG + (c % CG)];
    out_mat[i] = t;
  }
  if (threadIdx.x < TOK_L) {
    const int l = threadIdx.x, li = l >> 1;              // channel group 0 carries the per-pixel scalars
    float t = 0.f;
    for (int wp = 0; wp < NW; ++wp) t += svec[(wp * LP + li) * 2 + (l & 1)];
    out_vec[l] = t;
  }
}

// ---- forward, pass 1: chunk-local softmax statistics and un-normalised partial tokens ------------------------------
// grid (nchunks, 4*B).  part: [bk][chunk][ m[8] | s[8] | T[8][C] ]
// ITER = pixels per thread = chunk * LP / threads (compile time): the logits of phase 1 stay in registers for phase 2.
// All multiply-adds are packed FFMA2 (two channels per issue slot).
template <int CG, int ITER>
__global__ void __launch_bounds__(TOK_FWD_THREADS)
tok_fwd_chunk_kernel(const float* __restrict__ x, const float* __restrict__ wa, const float* __restrict__ ba,
                     float* __restrict__ part, TokGeom g) {
  constexpr int NT = TOK_FWD_THREADS, H2 = CG / 2;
  __shared__ float red[NT / 32][TOK_L];
  __shared__ float mfin[TOK_L];
  __shared__ uint64_t bar;
  extern __shared__ __align__(16) float dyn[];            // [chunk][C] staged rows | reduction scratch
  const int C = g.C, LP = 1 << g.lpshift;
  const int li = threadIdx.x & (LP - 1), tp = li & 3, cg = li >> 2;
  const int pin = threadIdx.x >> g.lpshift, ppi = NT >> g.lpshift;        // pixels per iteration
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = blockIdx.y;
  const int p0 = blockIdx.x * g.chunk;
  const int n = (int64_t)p0 + g.chunk < g.hw ? g.chunk : (int)(g.hw - p0);          // pixels of this chunk
  float* accs = dyn + g.chunk * C;
  stage_chunk(dyn, x + ((int64_t)bk * g.hw + p0) * C, (uint32_t)n * C * 4u, &bar);
  const float* xb = dyn + cg * CG;
  float2 w[2][H2];
  float b2[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    float r[CG];
    load_row<CG>(r, wa + (2 * tp + t) * C + cg * CG);
#pragma unroll
    for (int i = 0; i < H2; ++i) w[t][i] = make_float2(r[2 * i], r[2 * i + 1]);
    b2[t] = __ldg(ba + 2 * tp + t);
  }
  mbar_wait(&bar, 0);
  // phase 1: logits of the thread's pixels (kept) and the chunk's maximum per token
  float lg[ITER][2];
  float m[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int p = it * ppi + pin;
    float2 xp[H2];
    lds_pairs<CG>(xp, xb + (p < n ? p : 0) * C);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float2 d = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < H2; ++i) d = __ffma2_rn(w[t][i], xp[i], d);
      const float v = groups_sum(d.x + d.y, LP) + b2[t];
      lg[it][t] = p < n ? v : -INFINITY;                 // exp(-inf - m) = 0: dead pixels drop out of every sum
      m[t] = fmaxf(m[t], lg[it][t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float mm = pixels_max(m[t], LP);
    if (lane < 4) red[warp][2 * lane + t] = mm;          // lanes 0..3: cg = 0, tp = lane
  }
  __syncthreads();
  if (threadIdx.x < TOK_L) {
    float mm = red[0][threadIdx.x];
    for (int i = 1; i < NT / 32; ++i) mm = fmaxf(mm, red[i][threadIdx.x]);
    mfin[threadIdx.x] = mm;
  }
  __syncthreads();
  m[0] = mfin[2 * tp]; m[1] = mfin[2 * tp + 1];
  // phase 2: s[l] = sum exp(logit - m), T[l][c] = sum exp(logit - m) * x[c]   (x again from shared memory)
  float s[2] = {0.f, 0.f};
  float2 acc[2][H2];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < H2; ++i) acc[t][i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int p = it * ppi + pin;
    float2 xp[H2];
    lds_pairs<CG>(xp, xb + (p < n ? p : 0) * C);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const float e = __expf(lg[it][t] - m[t]);
      s[t] += e;
      const float2 e2 = make_float2(e, e);
#pragma unroll
      for (int i = 0; i < H2; ++i) acc[t][i] = __ffma2_rn(e2, xp[i], acc[t][i]);
    }
  }
  float tile[2][CG];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < H2; ++i) { tile[t][2 * i] = acc[t][i].x; tile[t][2 * i + 1] = acc[t][i].y; }
  float* out = part + ((int64_t)bk * g.nchunks + blockIdx.x) * (2 * TOK_L + TOK_L * C);
  reduce_tiles<CG, NT>(tile, s, accs, out + 2 * TOK_L, out + TOK_L, C, LP);
  if (threadIdx.x < TOK_L) out[threadIdx.x] = mfin[threadIdx.x];
}

// ---- forward, pass 2: combine the chunks of one (pair, frame) -------------------------------------------------------
// One warp per (pair-frame, token): lanes run over the chunks, every reduction is a fixed-order butterfly.
// grid = ceil(4*B*8 / 8) CTAs of 8 warps.  tokens: [bk][L][C]; stats: [bk][ M[8] | 1/S[8] ]
__global__ void __launch_bounds__(256)
tok_fwd_combine_kernel(const float* __restrict__ part, float* __restrict__ tokens, float* __restrict__ stats,
                       TokGeom g, int n_bk) {
  const int C = g.C, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5);            // = bk * 8 + l
  if (wid >= n_bk * TOK_L) return;
  const int bk = wid >> 3, l = wid & 7;
  const int stride = 2 * TOK_L + TOK_L * C;
  const float* pb = part + (int64_t)bk * g.nchunks * stride;
  float M = -INFINITY;
  for (int k = lane; k < g.nchunks; k += 32) M = fmaxf(M, pb[k * stride + l]);
  for (int d = 16; d > 0; d >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, d));
  float S = 0.f;
  for (int k = lane; k < g.nchunks; k += 32) S = fmaf(pb[k * stride + TOK_L + l], expf(pb[k * stride + l] - M), S);
  for (int d = 16; d > 0; d >>= 1) S += __shfl_xor_sync(0xffffffffu, S, d);
  const float rS = __fdiv_rn(1.f, S);
  for (int c0 = 0; c0 < C; c0 += 4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = lane; k < g.nchunks; k += 32) {
      const float sc = expf(pb[k * stride + l] - M);
      const float4 v = *reinterpret_cast<const float4*>(pb + k * stride + 2 * TOK_L + l * C + c0);
      t.x = fmaf(v.x, sc, t.x); t.y = fmaf(v.y, sc, t.y); t.z = fmaf(v.z, sc, t.z); t.w = fmaf(v.w, sc, t.w);
    }
    for (int d = 16; d > 0; d >>= 1) {
      t.x += __shfl_xor_sync(0xffffffffu, t.x, d); t.y += __shfl_xor_sync(0xffffffffu, t.y, d);
      t.z += __shfl_xor_sync(0xffffffffu, t.z, d); t.w += __shfl_xor_sync(0xffffffffu, t.w, d);
    }
    if (lane == 0)
      *reinterpret_cast<float4*>(tokens + (int64_t)bk * TOK_L * C + l * C + c0) =
          make_float4(__fmul_rn(t.x, rS), __fmul_rn(t.y, rS), __fmul_rn(t.z, rS), __fmul_rn(t.w, rS));
  }
  if (lane == 0) { stats[bk * 2 * TOK_L + l] = M; stats[bk * 2 * TOK_L + TOK_L + l] = rS; }
}

// ---- backward: one pass ---------------------------------------------------------------------------------------------
// grid (nchunks, 4*B).  part: [bk][chunk][ dW[8][C] | db[8] ]
template <int CG>
__global__ void __launch_bounds__(TOK_BWD_THREADS)
tok_bwd_chunk_kernel(const float* __restrict__ gtok, const float* __restrict__ x, const float* __restrict__ wa,
                     const float* __restrict__ ba, const float* __restrict__ tokens, const float* __restrict__ stats,
                     float* __restrict__ gx, float* __restrict__ part, TokGeom g) {
  constexpr int NT = TOK_BWD_THREADS;
  __shared__ float dsum[TOK_L];
  __shared__ uint64_t bar;
  extern __shared__ __align__(16) float dyn[];
  const int C = g.C, LP = 1 << g.lpshift;
  const int li = threadIdx.x & (LP - 1), tp = li & 3, cg = li >> 2;
  const int pin = threadIdx.x >> g.lpshift, ppi = NT >> g.lpshift;
  const int bk = blockIdx.y;
  const int p0 = blockIdx.x * g.chunk;
  const int n = (int64_t)p0 + g.chunk < g.hw ? g.chunk : (int)(g.hw - p0);
  float* accs = dyn + g.chunk * C;
  stage_chunk(dyn, x + ((int64_t)bk * g.hw + p0) * C, (uint32_t)n * C * 4u, &bar);
  const float* xb = dyn + cg * CG;
  const float* gt = gtok + (int64_t)bk * TOK_L * C;
  // D[l] = sum_p attn[l,p] * dattn[l,p] = <gtok[l,:], tokens[l,:]>
  if (threadIdx.x < TOK_L) {
    const float* tk = tokens + (int64_t)bk * TOK_L * C + threadIdx.x * C;
    float d = 0.f;
    for (int c = 0; c < C; ++c) d = fmaf(__ldg(gt + threadIdx.x * C + c), __ldg(tk + c), d);
    dsum[threadIdx.x] = d;
  }
  __syncthreads();
  constexpr int H2 = CG / 2;
  float2 w[2][H2], gv[2][H2], dw2[2][H2];
  float b2[2], M[2], rS[2], D[2], db[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int l = 2 * tp + t;
    float r[CG], q[CG];
    load_row<CG>(r, wa + l * C + cg * CG);
    load_row<CG>(q, gt + l * C + cg * CG);
#pragma unroll
    for (int i = 0; i < H2; ++i) {
      w[t][i] = make_float2(r[2 * i], r[2 * i + 1]);
      gv[t][i] = make_float2(q[2 * i], q[2 * i + 1]);
      dw2[t][i] = make_float2(0.f, 0.f);
    }
    b2[t] = __ldg(ba + l);
    M[t] = __ldg(stats + bk * 2 * TOK_L + l);
    rS[t] = __ldg(stats + bk * 2 * TOK_L + TOK_L + l);
    D[t] = dsum[l];
    db[t] = 0.f;
  }
  // after two exchange-and-halve steps over the 4 token-pair lanes, lane tp owns CG/4 channels of d x
  constexpr int Q = CG / 4;
  const int own = (tp & 1) * (CG / 2) + ((tp >> 1) & 1) * Q;
  float* gxb = gx + ((int64_t)bk * g.hw + p0) * C + cg * CG + own;
  mbar_wait(&bar, 0);
  for (int pb = 0; pb < n; pb += ppi) {
    const int p = pb + pin;
    const bool live = p < n;
    float2 xv[H2];
    lds_pairs<CG>(xv, xb + (live ? p : 0) * C);
    float2 a2[2], dl2[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      float2 d = make_float2(0.f, 0.f), e = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < H2; ++i) { d = __ffma2_rn(w[t][i], xv[i], d); e = __ffma2_rn(gv[t][i], xv[i], e); }
      const float lg = groups_sum(d.x + d.y, LP) + b2[t];
      const float da = groups_sum(e.x + e.y, LP);                      // d attn[l,p]
      const float a = live ? __fmul_rn(__expf(lg - M[t]), rS[t]) : 0.f;
      const float dl = __fmul_rn(a, da - D[t]);                        // d logit[l,p]
      db[t] += dl;
      a2[t] = make_float2(a, a);
      dl2[t] = make_float2(dl, dl);
    }
    float dx[CG];
#pragma unroll
    for (int i = 0; i < H2; ++i) {
      const float2 v = __ffma2_rn(a2[0], gv[0][i], __ffma2_rn(dl2[0], w[0][i], __ffma2_rn(a2[1], gv[1][i], __fmul2_rn(dl2[1], w[1][i]))));
      dx[2 * i] = v.x; dx[2 * i + 1] = v.y;
      dw2[0][i] = __ffma2_rn(dl2[0], xv[i], dw2[0][i]);
      dw2[1][i] = __ffma2_rn(dl2[1], xv[i], dw2[1][i]);
    }
    // sum d x over the 4 token-pair lanes: exchange halves with lane^1, then quarters with lane^2
    float h1[CG / 2], h2[Q];
#pragma unroll
    for (int i = 0; i < CG / 2; ++i) {
      const float send = (tp & 1) ? dx[i] : dx[i + CG / 2];
      const float keep = (tp & 1) ? dx[i + CG / 2] : dx[i];
      h1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
      const float send = (tp & 2) ? h1[i] : h1[i + Q];
      const float keep = (tp & 2) ? h1[i + Q] : h1[i];
      h2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    if (live) {
      float* o = gxb + (int64_t)p * C;
      if constexpr (Q == 4) *reinterpret_cast<float4*>(o) = make_float4(h2[0], h2[1], h2[2], h2[3]);
      else if constexpr (Q == 2) *reinterpret_cast<float2*>(o) = make_float2(h2[0], h2[1]);
      else o[0] = h2[0];
    }
  }
  float dw[2][CG];
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int i = 0; i < H2; ++i) { dw[t][2 * i] = dw2[t][i].x; dw[t][2 * i + 1] = dw2[t][i].y; }
  float* out = part + ((int64_t)bk * g.nchunks + blockIdx.x) * (TOK_L * C + TOK_L);
  reduce_tiles<CG, NT>(dw, db, accs, out, out + TOK_L * C, C, LP);
}

// gwa[L][C], gba[L] = sum over every (pair, frame, chunk) partial.  One CTA per output element: 256 threads stride over
// the partials, then a fixed-order butterfly + 8-warp sum (deterministic).  grid = L*C + L
__global__ void __launch_bounds__(256)
tok_bwd_combine_kernel(const float* __restrict__ part, float* __restrict__ gwa, float* __restrict__ gba, int n_part,
                       int C) {
  __shared__ float red[8];
  const int i = blockIdx.x, n = TOK_L * C + TOK_L;
  float t = 0.f;
  for (int k = threadIdx.x; k < n_part; k += 256) t += part[(int64_t)k * n + i];
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int wv = 0; wv < 8; ++wv) r += red[wv];
    if (i < TOK_L * C) gwa[i] = r; else gba[i - TOK_L * C] = r;
  }
}

static bool tok_use_mma(int C) { return option(OPT_TOK_VARIANT) != 0 && tok_mma_supported(C); }
static int tok_geom(TokGeom& g, int B, int C, int64_t hw, const void* x, int dtype, int layout) {
  if (B <= 0 || C <= 0 || hw <= 0 || !x) return fail(SMOW_EINVAL, "tokenizer: bad shape / null pointer");
  if (dtype != SMOW_F32 || layout != SMOW_NDHWC)
    return fail(SMOW_EDTYPE, "tokenizer: built for fp32 channels_last_3d (NDHWC) stacks only");
  const int CG = C < 16 ? C : 16;
  g.C = C; g.hw = hw; g.G = C / CG; g.lpshift = -1;
  for (int s = 2; s < 6; ++s) if ((4 << (s - 2)) == 4 * g.G) g.lpshift = s;
  if ((CG != 4 && CG != 8 && CG != 16) || C % CG || g.lpshift < 0)
    return fail(SMOW_EINVAL, "tokenizer: C must be 4, 8 or 16 * (1, 2, 4, 8) (got C = %d)", C);
  if (!aligned16(x)) return fail(SMOW_EALIGN, "tokenizer: stack not 16 B aligned");
  g.chunk = tok_use_mma(C) ? tok_mma_chunk_px(C) : tok_chunk_px(C);
  g.nchunks = (int)((hw + g.chunk - 1) / g.chunk);
  if ((int64_t)4 * B > 65535) return fail(SMOW_ERANGE, "tokenizer: batch too large for one launch");
  return 0;
}
static size_t tok_smem(const TokGeom& g, int threads) {
  const int LP = 1 << g.lpshift, CG = g.C / g.G, NW = threads / 32;
  return (size_t)((size_t)g.chunk * g.C + NW * LP * 2 * CG + NW * LP * 2) * sizeof(float);
}
template <typename K> static void tok_allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace smow

using namespace smow;

extern "C" {

int64_t smow_tokenizer_workspace_bytes(int B, int C, int64_t hw) {
  if (C <= 0 || hw <= 0 || B <= 0) return 0;
  // sized for the smaller chunk of the two kernel families, so that the knob can change between the query and the call
  int64_t chunk = tok_chunk_px(C);
  if (tok_mma_supported(C) && tok_mma_chunk_px(C) < chunk) chunk = tok_mma_chunk_px(C);
  const int64_t nchunks = (hw + chunk - 1) / chunk;
  return (int64_t)4 * B * nchunks * (2 * TOK_L + TOK_L * C) * (int64_t)sizeof(float);
}

int smow_tokenizer_fwd(const void* x, const float* wa, const float* ba, float* tokens, float* stats, int B, int C,
                       int64_t hw, int dtype, int layout, void* ws, int64_t ws_bytes, void* stream) {
  TokGeom g;
  if (int rc = tok_geom(g, B, C, hw, x, dtype, layout)) return rc;
  if (!wa || !ba || !tokens || !stats) return fail(SMOW_EINVAL, "tokenizer: null pointer");
  if (!ws || ws_bytes < smow_tokenizer_workspace_bytes(B, C, hw) || !aligned16(ws))
    return fail(SMOW_EINVAL, "tokenizer: workspace of smow_tokenizer_workspace_bytes() bytes required");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(ws);
  const dim3 grid(g.nchunks, 4 * B);
  const size_t smem = tok_smem(g, TOK_FWD_THREADS);
  // C = 16 / 32 (the two models): the GEMMs run as 3xTF32 tensor-core MMAs (tokenizer_mma.cu) unless tok_variant = 0
  const bool mma = tok_use_mma(C);
#define SMOW_TOK_FWD(CG, ITER)                                \
  tok_allow_smem(tok_fwd_chunk_kernel<CG, ITER>, smem);       \
  tok_fwd_chunk_kernel<CG, ITER><<<grid, TOK_FWD_THREADS, smem, st>>>((const float*)x, wa, ba, part, g)
  // pixels per thread = chunk * LP / threads: 8 for C <= 16 (512-pixel chunks, 4 lanes per pixel), 16 beyond
  if (mma) tok_fwd_mma_launch((const float*)x, nullptr, wa, ba, part, g, B, st);
  else switch (C / g.G) {
    case 4: SMOW_TOK_FWD(4, 8); break;
    case 8: SMOW_TOK_FWD(8, 8); break;
    default: if (g.G == 1) { SMOW_TOK_FWD(16, 8); } else { SMOW_TOK_FWD(16, 16); } break;
  }
#undef SMOW_TOK_FWD
  tok_fwd_combine_kernel<<<(4 * B * TOK_L + 7) / 8, 256, 0, st>>>(part, tokens, stats, g, 4 * B);
  count_launch(2);
  return check_launch("tokenizer_fwd");
}

int smow_tokenizer_bwd(const float* gtokens, const void* x, const float* wa, const float* ba, const float* tokens,
                       const float* stats, void* gx, float* gwa, float* gba, int B, int C, int64_t hw, int dtype,
                       int layout, void* ws, int64_t ws_bytes, void* stream) {
  TokGeom g;
  if (int rc = tok_geom(g, B, C, hw, x, dtype, layout)) return rc;
  if (!gtokens || !wa || !ba || !tokens || !stats || !gx || !gwa || !gba) return fail(SMOW_EINVAL, "tokenizer: null pointer");
  if (!ws || ws_bytes < smow_tokenizer_workspace_bytes(B, C, hw) || !aligned16(ws) || !aligned16(gx))
    return fail(SMOW_EINVAL, "tokenizer: workspace of smow_tokenizer_workspace_bytes() bytes required");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(ws);
  const dim3 grid(g.nchunks, 4 * B);
  const size_t smem = tok_smem(g, TOK_BWD_THREADS);
  const bool mma = tok_use_mma(C);
#define SMOW_TOK_BWD(CG)                                 \
  tok_allow_smem(tok_bwd_chunk_kernel<CG>, smem);        \
  tok_bwd_chunk_kernel<CG><<<grid, TOK_BWD_THREADS, smem, st>>>(gtokens, (const float*)x, wa, ba, tokens, stats, (float*)gx, part, g)
  if (mma) tok_bwd_mma_launch(gtokens, (const float*)x, nullptr, wa, ba, tokens, stats, (float*)gx, part, g, B, st);
  else switch (C / g.G) {
    case 4: SMOW_TOK_BWD(4); break;
    case 8: SMOW_TOK_BWD(8); break;
    default: SMOW_TOK_BWD(16); break;
  }
#undef SMOW_TOK_BWD
  tok_bwd_combine_kernel<<<TOK_L * C + TOK_L, 256, 0, st>>>(part, gwa, gba, 4 * B * g.nchunks, C);
  count_launch(2);
  return check_launch("tokenizer_bwd");
}

// ---- rows A1 + N2 fused: OFW.flow_warp -> Transformer_Encoder's pooling without the stack in HBM -------------------------
static int warp_tok_geom(TokGeom& g, TokWarpSrc& src, const void* x, const float* flow, const float* xs, const float* ys,
                         int B, int C, int H, int W, int dtype, int layout) {
  if (H <= 0 || W <= 0) return fail(SMOW_EINVAL, "warp_tokenizer: bad shape");
  if (int rc = tok_geom(g, B, C, (int64_t)H * W, x, dtype, layout)) return rc;
  if (!flow || !xs || !ys) return fail(SMOW_EINVAL, "warp_tokenizer: null pointer");
  if (!tok_use_mma(C))
    return fail(SMOW_EDTYPE, "warp_tokenizer: built on the tensor-core tokenizer kernels (C = 16 / 32, tok_variant != 0)");
  if ((int64_t)B * 4 * H * W * C >= ((int64_t)1 << 31) || (int64_t)H * W >= ((int64_t)1 << 29))
    return fail(SMOW_ERANGE, "warp_tokenizer: tensor too large for 32-bit tile indexing");
  src.x = (const float*)x; src.flow = flow; src.xs = xs; src.ys = ys; src.H = H; src.W = W; src.wshift = -1;
  for (int s = 0; s < 31; ++s) if ((1 << s) == W) src.wshift = s;
  return 0;
}

int smow_warp_tokenizer_supported(int C) { return tok_use_mma(C) ? 1 : 0; }

int smow_warp_tokenizer_fwd(const void* x, const float* flow, const float* xs, const float* ys, const float* wa,
                            const float* ba, float* tokens, float* stats, int B, int C, int H, int W, int dtype,
                            int layout, void* ws, int64_t ws_bytes, void* stream) {
  TokGeom g;
  TokWarpSrc src;
  if (int rc = warp_tok_geom(g, src, x, flow, xs, ys, B, C, H, W, dtype, layout)) return rc;
  if (!wa || !ba || !tokens || !stats) return fail(SMOW_EINVAL, "warp_tokenizer: null pointer");
  if (!ws || ws_bytes < smow_tokenizer_workspace_bytes(B, C, g.hw) || !aligned16(ws))
    return fail(SMOW_EINVAL, "warp_tokenizer: workspace of smow_tokenizer_workspace_bytes() bytes required");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(ws);
  tok_fwd_mma_launch(nullptr, &src, wa, ba, part, g, B, st);
  tok_fwd_combine_kernel<<<(4 * B * TOK_L + 7) / 8, 256, 0, st>>>(part, tokens, stats, g, 4 * B);
  count_launch(2);
  return check_launch("warp_tokenizer_fwd");
}

int smow_warp_tokenizer_bwd(const float* gtokens, const void* x, const float* flow, const float* xs, const float* ys,
                            const float* wa, const float* ba, const float* tokens, const float* stats, void* gstack,
                            float* gwa, float* gba, int B, int C, int H, int W, int dtype, int layout, void* ws,
                            int64_t ws_bytes, void* stream) {
  TokGeom g;
  TokWarpSrc src;
  if (int rc = warp_tok_geom(g, src, x, flow, xs, ys, B, C, H, W, dtype, layout)) return rc;
  if (!gtokens || !wa || !ba || !tokens || !stats || !gstack || !gwa || !gba) return fail(SMOW_EINVAL, "warp_tokenizer: null pointer");
  if (!ws || ws_bytes < smow_tokenizer_workspace_bytes(B, C, g.hw) || !aligned16(ws) || !aligned16(gstack))
    return fail(SMOW_EINVAL, "warp_tokenizer: workspace of smow_tokenizer_workspace_bytes() bytes required");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(ws);
  tok_bwd_mma_launch(gtokens, nullptr, &src, wa, ba, tokens, stats, (float*)gstack, part, g, B, st);
  tok_bwd_combine_kernel<<<TOK_L * C + TOK_L, 256, 0, st>>>(part, gwa, gba, 4 * B * g.nchunks, C);
  count_launch(2);
  return check_launch("warp_tokenizer_bwd");
}

}  // extern "C"
