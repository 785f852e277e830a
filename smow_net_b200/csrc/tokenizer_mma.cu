// Row N2 on the tensor cores: the tokenizer's three small GEMMs (8 tokens wide) as warp-level TF32 MMAs with the
// 3xTF32 split (hi + lo parts of both operands, fp32 accumulation), which keeps fp32-level accuracy (the 1e-5 parity bar of
// tests/) while taking ~4/5 of the instructions off the FP32 pipe that bounds tokenizer.cu (profiles/r2_notes.md).
// Reference: models/SMOW_Net.py:171-190, models/SMOW_Net_LW.py:190-209.
//
// Shapes are far below a tcgen05 tile (N = 8 tokens, K = 16 / 32 channels), so the warp-level m16n8k8 MMA is the fitting
// instruction; the kernels stay bandwidth-shaped: one bulk copy stages a chunk of pixel rows, every warp works on its own
// pixels out of shared memory, per-chunk partials leave the CTA (same layout as tokenizer.cu: the combine kernels are shared).
//
// Fragment algebra (lane = 4*g + t).  All index permutations below are free because a GEMM does not care how its K (and M, N)
// indices are numbered as long as both operands agree:
//   GEMM1  logit^T[l, p] = sum_c W[l,c] X[p,c]        A = [W_hi ; W_lo] (rows 0-7 / 8-15: one MMA yields both partial
//          products, logit = row g + row g+8), B = X[p = g][c]: thread (g,t) feeds the float4 X[p_g][16q+4t ..+3]
//          (conflict-free LDS.128) as k-steps 2q, 2q+1.  D: thread holds logit[l = g][p = 2t, 2t+1].
//   GEMM2  T^T[c, l] = sum_p X[p,c] E[p,l]            (tokens / dW) B = E straight from GEMM1's accumulator layout
//          (k = t <-> p = 2t, k = t+4 <-> p = 2t+1), A = X[p][16j + 2g, +1] as rows g / g+8.
//   GEMM3  dX[p, c] = sum_l' S[p,l'] V[l',c]          (backward) S = [attn | dlogit] transposed across the warp with 4
//          shuffles per value pair, V = [gtok ; W] pre-split in shared memory in fragment order.
#include "bulk.cuh"
#include "tokenizer.cuh"

namespace smow {

constexpr int TOKM_FWD_THREADS = 256, TOKM_BWD_THREADS = 128;

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// x = hi + lo + O(2^-22 x): hi, lo both exactly representable in TF32.  Round-to-nearest conversions (cvt.rna is emulated
// with ~4 integer instructions on sm_100a): used for the operands that are split once per CTA.
__device__ __forceinline__ void split_tf32_rn(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// The per-pixel operands are split with two instructions instead: hi = x with the 13 low mantissa bits cleared, lo = x - hi
// (exact in fp32); the tensor core ignores the low 13 bits of lo, an error below 2^-21 |x| — still fp32-class accuracy.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// A operand of GEMM1 for the 8 x C matrix `src` (row l = g): per k-step {hi(k=t), lo(k=t), hi(k=t+4), lo(k=t+4)}
template <int C>
__device__ __forceinline__ void load_row_frags(uint32_t (&f)[C / 8][4], const float* __restrict__ src, int g, int t) {
#pragma unroll
  for (int q = 0; q < C / 16; ++q) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + g * C + 16 * q + 4 * t));
    split_tf32_rn(v.x, f[2 * q][0], f[2 * q][1]);
    split_tf32_rn(v.y, f[2 * q][2], f[2 * q][3]);
    split_tf32_rn(v.z, f[2 * q + 1][0], f[2 * q + 1][1]);
    split_tf32_rn(v.w, f[2 * q + 1][2], f[2 * q + 1][3]);
  }
}

// zero the rows of the staged chunk that the bulk copy does not write (ragged last chunk): dead pixels then contribute
// exact zeros to every MMA instead of whatever the shared memory held
__device__ __forceinline__ void zero_dead_rows(float* xs, int n, int chunk, int C, int nthreads) {
  for (int i = n * C + threadIdx.x; i < chunk * C; i += nthreads) xs[i] = 0.f;
}

// ---- fused staging (rows A1 + N2) -------------------------------------------------------------------------------------------
// Frame f of pair b as the kernels see it: f = 0 / 3 are the input frames themselves, f = 1 / 2 their warps (reference
// models/SMOW_Net.py:634-636).  stage_warped_rows writes the warped rows [p0, p0 + n) of frame 1 + t into shared memory —
// what warp_fwd_ndhwc_shfl_kernel would have written to HBM, bit for bit: a warp owns 32 consecutive pixels, lane l runs the
// reference's coordinate chain for pixel l once, then in Q steps every lane serves one (pixel, 4-channel vector) with the
// footprint fetched by indexed shuffles: 4 tap gathers, FMAs in ATen's order (nw, ne, sw, se; out-of-bounds taps skipped).
// Rows >= n (ragged last chunk) are written as zeros.
template <int C, int CHUNK, int NT>
__device__ __forceinline__ void stage_warped_rows(float* __restrict__ rows, const TokWarpSrc& s, int b, int t, int p0, int n) {
  constexpr int Q = C / 4, QSHIFT = Q == 4 ? 2 : 3, PPW = 32 / Q, GPW = (CHUNK / 32) / (NT / 32);
  static_assert(Q == 4 || Q == 8, "C = 16 / 32");
  static_assert(GPW * (NT / 32) * 32 == CHUNK, "whole 32-pixel groups per warp");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = s.H * s.W, W = s.W;
  const int v = lane & (Q - 1), sub = lane >> QSHIFT;
  const float* src = s.x + (int64_t)(b * 2 + t) * HW * C;      // CTA-uniform base + 32-bit element offsets below
  const float* fl = s.flow + ((int64_t)(b * 2) * 2 + t) * HW;
  // The staging is bound by instruction issue and load latency, not by bandwidth, so it is written for few instructions and
  // many independent loads: the warp's GPW groups are independent (all flow loads first, then the chains, then the gathers);
  // a tap outside the image gets weight 0 and is re-pointed at the nw tap (always in bounds), which leaves the sum bit for
  // bit what ATen's skipped tap gives (r + c * 0 = r) with four unconditional loads and a flat FMA chain; a dead pixel (ragged
  // chunk) gets four zero weights.  Per pixel one packed word (nw offset, x / y step present) and four weights are shuffled.
  float fx[GPW], fy[GPW];
#pragma unroll
  for (int k = 0; k < GPW; ++k) {
    const int pl = (warp * GPW + k) * 32 + lane;
    const bool live = pl < n;
    fx[k] = live ? __ldg(fl + p0 + pl) : 0.f;
    fy[k] = live ? __ldg(fl + p0 + pl + 2 * (int64_t)HW) : 0.f;
  }
  int own_o[GPW];
  float own_nw[GPW], own_ne[GPW], own_sw[GPW], own_se[GPW];
#pragma unroll
  for (int k = 0; k < GPW; ++k) {
    const int pl = (warp * GPW + k) * 32 + lane;
    const bool live = pl < n;
    const int p = live ? p0 + pl : p0;
    const int h = s.wshift >= 0 ? (p >> s.wshift) : (p / W), w = p - h * W;
    const Footprint fp = footprint_auto(__ldg(s.xs + w), __ldg(s.ys + h), fx[k], fy[k], W, s.H);
    own_nw[k] = live ? __fmul_rn(fp.wx0, fp.wy0) : 0.f;
    own_ne[k] = live && fp.x1ok ? __fmul_rn(fp.wx1, fp.wy0) : 0.f;
    own_sw[k] = live && fp.y1ok ? __fmul_rn(fp.wx0, fp.wy1) : 0.f;
    own_se[k] = live && fp.x1ok && fp.y1ok ? __fmul_rn(fp.wx1, fp.wy1) : 0.f;
    own_o[k] = ((fp.y0 * W + fp.x0) << 2) | (fp.x1ok ? 1 : 0) | (fp.y1ok ? 2 : 0);     // HW < 2^29 (checked by the host)
  }
  const int rowC = W * C;
#pragma unroll
  for (int k = 0; k < GPW; ++k) {
    const int g0 = (warp * GPW + k) * 32;
    float4 r[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      const int sl = j * PPW + sub;                               // lane that holds this pixel's footprint
      const int o = __shfl_sync(0xffffffffu, own_o[k], sl);
      const float nw = __shfl_sync(0xffffffffu, own_nw[k], sl), ne = __shfl_sync(0xffffffffu, own_ne[k], sl);
      const float sw = __shfl_sync(0xffffffffu, own_sw[k], sl), se = __shfl_sync(0xffffffffu, own_se[k], sl);
      const uint32_t o0 = (uint32_t)(o >> 2) * C + v * 4, dx = (o & 1) ? C : 0, dy = (o & 2) ? rowC : 0;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + o0));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(src + o0 + dx));
      const float4 c2 = __ldg(reinterpret_cast<const float4*>(src + o0 + dy));
      const float4 c3 = __ldg(reinterpret_cast<const float4*>(src + o0 + dx + dy));
      float4 q;                                                   // ATen's order: nw, ne, sw, se
      q.x = __fmul_rn(a.x, nw); q.y = __fmul_rn(a.y, nw); q.z = __fmul_rn(a.z, nw); q.w = __fmul_rn(a.w, nw);
      q.x = fmaf(c1.x, ne, q.x); q.y = fmaf(c1.y, ne, q.y); q.z = fmaf(c1.z, ne, q.z); q.w = fmaf(c1.w, ne, q.w);
      q.x = fmaf(c2.x, sw, q.x); q.y = fmaf(c2.y, sw, q.y); q.z = fmaf(c2.z, sw, q.z); q.w = fmaf(c2.w, sw, q.w);
      q.x = fmaf(c3.x, se, q.x); q.y = fmaf(c3.y, se, q.y); q.z = fmaf(c3.z, se, q.z); q.w = fmaf(c3.w, se, q.w);
      r[j] = g0 + sl < n ? q : make_float4(0.f, 0.f, 0.f, 0.f);   // dead rows are exact zeros whatever x holds
    }
#pragma unroll
    for (int j = 0; j < Q; ++j) *reinterpret_cast<float4*>(rows + (g0 + j * PPW + sub) * C + v * 4) = r[j];
  }
}

// Stage the chunk [p0, p0 + n) of stack frame bk = 4 b + f into `rows`.  Plain form: one bulk copy out of the stack.  Fused
// form: bulk copy out of input frame 0 / 1 for f = 0 / 3, warp for f = 1 / 2.  Returns true when the caller has to wait on
// `bar` (a bulk copy is in flight) after the CTA-wide barrier both forms need (mbarrier initialised / rows written).
template <int C, int CHUNK, int NT, bool FUSED>
__device__ __forceinline__ bool stage_chunk(float* rows, uint64_t* bar, const float* __restrict__ x, const TokWarpSrc& s,
                                            int bk, int p0, int n, int64_t hw) {
  const int f = bk & 3;
  const bool warped = FUSED && (f == 1 || f == 2);
  // the warp staging comes first: every thread is still converged here, so its shuffles need no re-convergence code
  if (warped) stage_warped_rows<C, CHUNK, NT>(rows, s, bk >> 2, f - 1, p0, n);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    if (!warped) {
      const float* from = FUSED ? s.x + ((int64_t)((bk >> 2) * 2 + (f == 3 ? 1 : 0)) * hw + p0) * C
                                : x + ((int64_t)bk * hw + p0) * C;
      mbar_expect_tx(bar, (uint32_t)n * C * 4u);
      bulk_g2s(rows, from, (uint32_t)n * C * 4u, bar);
    }
  }
  if (!warped && n < CHUNK) zero_dead_rows(rows, n, CHUNK, C, NT);
  return !warped;
}

// ---- forward -----------------------------------------------------------------------------------------------------------
// grid (nchunks, 4*B), 8 warps, one chunk per CTA (several CTAs per SM overlap each other's staging: a persistent,
// double-buffered variant with fewer resident warps measured 10-20 % slower — the kernels are bound by the latency of their
// dependent MMA / shared-memory chains, not by the copy).  Warp w owns pixels [w*PW, (w+1)*PW) of the chunk, PW = CHUNK / 8,
// in groups of 8.  Softmax statistics are warp-local (max over the warp's pixels), merged once per CTA.
template <int C, int CHUNK, bool FUSED>
__global__ void __launch_bounds__(TOKM_FWD_THREADS)
tok_fwd_mma_kernel(const float* __restrict__ x, const TokWarpSrc src, const float* __restrict__ wa,
                   const float* __restrict__ ba, float* __restrict__ part, TokGeom geo) {
  constexpr int NT = TOKM_FWD_THREADS, NW = NT / 32, PW = CHUNK / NW, NG = PW / 8;
  __shared__ uint64_t bar;
  extern __shared__ __align__(16) float dyn[];            // [CHUNK][C] staged rows | per-warp partials
  float* xs = dyn;
  float* tw = dyn + CHUNK * C;                            // [NW][8][C]
  float* mw = tw + NW * TOK_L * C;                        // [NW][8]
  float* sw = mw + NW * TOK_L;                            // [NW][8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // fused form: frame order 1, 0, 2, 3 within a pair — the CTAs that warp input frame t and the ones that copy it run back
  // to back, so the second reader finds the frame in the L2 (ncu, B = 64, C = 32, dispatched "all warps first": 544 MB read
  // from HBM for a 268 MB x)
  const int bk = FUSED ? (int)(blockIdx.y & ~3u) + ((0x3201 >> (4 * (blockIdx.y & 3))) & 3) : (int)blockIdx.y;
  const int p0 = blockIdx.x * CHUNK;
  const int n = (int64_t)p0 + CHUNK < geo.hw ? CHUNK : (int)(geo.hw - p0);
  const bool copied = stage_chunk<C, CHUNK, NT, FUSED>(dyn, &bar, x, src, bk, p0, n, geo.hw);
  uint32_t wf[C / 8][4];
  load_row_frags<C>(wf, wa, g, t);
  const float bias = __ldg(ba + g);
  __syncthreads();                                        // barrier initialised, dead rows zeroed / warped rows written
  if (copied) mbar_wait(&bar, 0);
  const int wbase = warp * PW;
  {
    // phase 1: logits of the warp's pixels (kept in registers) and their maximum per token
    float lg[NG][2];
    float mx = -INFINITY;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const float* xrow = xs + (wbase + gi * 8 + g) * C + 4 * t;
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < C / 16; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(xrow + 16 * q);
        uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
        split_tf32(v.x, h0, l0); split_tf32(v.y, h1, l1); split_tf32(v.z, h2, l2); split_tf32(v.w, h3, l3);
        mma_tf32(d, wf[2 * q][0], wf[2 * q][1], wf[2 * q][2], wf[2 * q][3], h0, h1);
        mma_tf32(d, wf[2 * q][0], wf[2 * q][1], wf[2 * q][2], wf[2 * q][3], l0, l1);
        mma_tf32(d, wf[2 * q + 1][0], wf[2 * q + 1][1], wf[2 * q + 1][2], wf[2 * q + 1][3], h2, h3);
        mma_tf32(d, wf[2 * q + 1][0], wf[2 * q + 1][1], wf[2 * q + 1][2], wf[2 * q + 1][3], l2, l3);
      }
      const int p = wbase + gi * 8 + 2 * t;
      lg[gi][0] = p < n ? d[0] + d[2] + bias : -INFINITY;     // exp(-inf - m) = 0: dead pixels drop out of every sum
      lg[gi][1] = p + 1 < n ? d[1] + d[3] + bias : -INFINITY;
      mx = fmaxf(mx, fmaxf(lg[gi][0], lg[gi][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float mm = mx == -INFINITY ? 0.f : mx;              // a warp without live pixels: every exponential is 0
    // phase 2: s[l] = sum exp(logit - m), T^T[c][l] = sum x[c] exp(logit - m); two accumulator sets (even / odd groups)
    float s_sum = 0.f;
    float acc[2][C / 16][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int j = 0; j < C / 16; ++j) { acc[u][j][0] = 0.f; acc[u][j][1] = 0.f; acc[u][j][2] = 0.f; acc[u][j][3] = 0.f; }
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const float e0 = __expf(lg[gi][0] - mm), e1 = __expf(lg[gi][1] - mm);
      s_sum += e0 + e1;
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(e0, bh0, bl0); split_tf32(e1, bh1, bl1);
      const float* xp = xs + (wbase + gi * 8 + 2 * t) * C + 2 * g;
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {
        const float2 u0 = *reinterpret_cast<const float2*>(xp + 16 * j);
        const float2 u1 = *reinterpret_cast<const float2*>(xp + C + 16 * j);
        uint32_t ah[4], al[4];
        split_tf32(u0.x, ah[0], al[0]); split_tf32(u0.y, ah[1], al[1]);
        split_tf32(u1.x, ah[2], al[2]); split_tf32(u1.y, ah[3], al[3]);
        mma_tf32(acc[gi & 1][j], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
        mma_tf32(acc[gi & 1][j], al[0], al[1], al[2], al[3], bh0, bh1);
        mma_tf32(acc[gi & 1][j], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
      }
    }
    s_sum += __shfl_xor_sync(0xffffffffu, s_sum, 1);
    s_sum += __shfl_xor_sync(0xffffffffu, s_sum, 2);
    // acc[j]: {T[2t][16j+2g], T[2t+1][16j+2g], T[2t][16j+2g+1], T[2t+1][16j+2g+1]}
    float* twp = tw + warp * TOK_L * C;
#pragma unroll
    for (int j = 0; j < C / 16; ++j) {
      *reinterpret_cast<float2*>(twp + (2 * t) * C + 16 * j + 2 * g) =
          make_float2(acc[0][j][0] + acc[1][j][0], acc[0][j][2] + acc[1][j][2]);
      *reinterpret_cast<float2*>(twp + (2 * t + 1) * C + 16 * j + 2 * g) =
          make_float2(acc[0][j][1] + acc[1][j][1], acc[0][j][3] + acc[1][j][3]);
    }
    if (t == 0) { mw[warp * TOK_L + g] = mx; sw[warp * TOK_L + g] = s_sum; }
    __syncthreads();
    // merge the 8 warps in a fixed order
    float* out = part + ((int64_t)bk * geo.nchunks + blockIdx.x) * (2 * TOK_L + TOK_L * C);
    for (int i = threadIdx.x; i < TOK_L * C + TOK_L; i += NT) {
      const bool is_s = i >= TOK_L * C;
      const int l = is_s ? i - TOK_L * C : i / C;
      float m = mw[l];
#pragma unroll
      for (int w = 1; w < NW; ++w) m = fmaxf(m, mw[w * TOK_L + l]);
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const float sc = __expf(mw[w * TOK_L + l] - m);         // exp(-inf) = 0 for a warp without live pixels
        v = fmaf(is_s ? sw[w * TOK_L + l] : tw[w * TOK_L * C + i], sc, v);
      }
      if (is_s) { out[l] = m; out[TOK_L + l] = v; }
      else out[2 * TOK_L + i] = v;
    }
  }
}

// ---- backward ----------------------------------------------------------------------------------------------------------
// grid (nchunks, 4*B), 4 warps; warp w owns pixels [w*PW, (w+1)*PW), 16 per iteration (two 8-pixel groups A, B).
template <int C, int CHUNK, bool FUSED>
__global__ void __launch_bounds__(TOKM_BWD_THREADS, C == 16 ? 5 : 3)
tok_bwd_mma_kernel(const float* __restrict__ gtok, const float* __restrict__ x, const TokWarpSrc src,
                   const float* __restrict__ wa,
                   const float* __restrict__ ba, const float* __restrict__ tokens, const float* __restrict__ stats,
                   float* __restrict__ gx, float* __restrict__ part, TokGeom geo) {
  constexpr int NT = TOKM_BWD_THREADS, NW = NT / 32, PW = CHUNK / NW, NI = PW / 16;
  __shared__ float dsum[TOK_L];
  __shared__ uint64_t bar;
  extern __shared__ __align__(16) float dyn[];            // [CHUNK][C] staged rows | V fragments | per-warp partials
  const float* xs = dyn;
  uint4* vs = reinterpret_cast<uint4*>(dyn + CHUNK * C);  // [C/8 n-tiles][2 k-steps][32 lanes] {hi b0, hi b1, lo b0, lo b1}
  float* dww = dyn + CHUNK * C + (C / 8) * 2 * 32 * 4;    // [NW][8][C]
  float* dbw = dww + NW * TOK_L * C;                      // [NW][8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // fused form: frame order 1, 0, 2, 3 within a pair — the CTAs that warp input frame t and the ones that copy it run back
  // to back, so the second reader finds the frame in the L2 (ncu, B = 64, C = 32, dispatched "all warps first": 544 MB read
  // from HBM for a 268 MB x)
  const int bk = FUSED ? (int)(blockIdx.y & ~3u) + ((0x3201 >> (4 * (blockIdx.y & 3))) & 3) : (int)blockIdx.y;
  const int p0 = blockIdx.x * CHUNK;
  const int n = (int64_t)p0 + CHUNK < geo.hw ? CHUNK : (int)(geo.hw - p0);
  const bool copied = stage_chunk<C, CHUNK, NT, FUSED>(dyn, &bar, x, src, bk, p0, n, geo.hw);
  // B operand of GEMM3 in fragment order: n-tile j, column n = g <-> channel 16(j>>1) + 4(g>>1) + 2(j&1) + (g&1), so that a
  // thread's accumulators of n-tiles 2q, 2q+1 are the 4 consecutive channels 16q + 4t .. +3 of its pixel.
  // k-step 0: attn x gtok (per pair-frame), k-step 1: dlogit x W (constant)
  auto fill_v = [&](const float* __restrict__ src, int ks) {
    for (int i = threadIdx.x; i < (C / 8) * 32; i += NT) {
      const int ln = i & 31, j = i >> 5;
      const int gg = ln >> 2, tt = ln & 3;
      const int ch = 16 * (j >> 1) + 4 * (gg >> 1) + 2 * (j & 1) + (gg & 1);
      uint4 v;
      split_tf32_rn(__ldg(src + (2 * tt) * C + ch), v.x, v.z);
      split_tf32_rn(__ldg(src + (2 * tt + 1) * C + ch), v.y, v.w);
      vs[(j * 2 + ks) * 32 + ln] = v;
    }
  };
  const float* gt = gtok + (int64_t)bk * TOK_L * C;
  fill_v(gt, 0);
  fill_v(wa, 1);
  // D[l] = sum_p attn[l,p] * dattn[l,p] = <gtok[l,:], tokens[l,:]>: one warp, 4 lanes per token
  if (warp == 0) {
    const float* tk = tokens + (int64_t)bk * TOK_L * C + g * C;
    float d = 0.f;
    for (int c = t; c < C; c += 4) d = fmaf(__ldg(gt + g * C + c), __ldg(tk + c), d);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    if (t == 0) dsum[g] = d;
  }
  uint32_t wf[C / 8][4], gf[C / 8][4];
  load_row_frags<C>(wf, wa, g, t);
  load_row_frags<C>(gf, gt, g, t);
  const float bias = __ldg(ba + g);
  const float M = __ldg(stats + bk * 2 * TOK_L + g), rS = __ldg(stats + bk * 2 * TOK_L + TOK_L + g);
  const int src0 = 8 * t + (g >> 1);                      // lane holding (l = 2t, pixel g) after GEMM1; l = 2t+1: +4
  const bool odd = g & 1;
  __syncthreads();                                        // barrier initialised, dead rows zeroed, V fragments and D ready
  const float D = dsum[g];
  if (copied) mbar_wait(&bar, 0);
  {
    float db = 0.f;
    float dwt[C / 16][4];
#pragma unroll
    for (int j = 0; j < C / 16; ++j) { dwt[j][0] = 0.f; dwt[j][1] = 0.f; dwt[j][2] = 0.f; dwt[j][3] = 0.f; }
    float* gxb = gx + ((int64_t)bk * geo.hw + p0) * C + 4 * t;
#pragma unroll 1
    for (int it = 0; it < NI; ++it) {
      const int base = warp * PW + it * 16;
      if (base >= n) break;                               // warp-uniform: nothing live from here on
      float at[2][2], dlt[2][2];                          // transposed: [group][l = 2t, 2t+1] of pixel g
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float* xrow = xs + (base + 8 * h + g) * C + 4 * t;
        float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < C / 16; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(xrow + 16 * q);
          uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
          split_tf32(v.x, h0, l0); split_tf32(v.y, h1, l1); split_tf32(v.z, h2, l2); split_tf32(v.w, h3, l3);
          mma_tf32(d1, wf[2 * q][0], wf[2 * q][1], wf[2 * q][2], wf[2 * q][3], h0, h1);
          mma_tf32(d2, gf[2 * q][0], gf[2 * q][1], gf[2 * q][2], gf[2 * q][3], h0, h1);
          mma_tf32(d1, wf[2 * q][0], wf[2 * q][1], wf[2 * q][2], wf[2 * q][3], l0, l1);
          mma_tf32(d2, gf[2 * q][0], gf[2 * q][1], gf[2 * q][2], gf[2 * q][3], l0, l1);
          mma_tf32(d1, wf[2 * q + 1][0], wf[2 * q + 1][1], wf[2 * q + 1][2], wf[2 * q + 1][3], h2, h3);
          mma_tf32(d2, gf[2 * q + 1][0], gf[2 * q + 1][1], gf[2 * q + 1][2], gf[2 * q + 1][3], h2, h3);
          mma_tf32(d1, wf[2 * q + 1][0], wf[2 * q + 1][1], wf[2 * q + 1][2], wf[2 * q + 1][3], l2, l3);
          mma_tf32(d2, gf[2 * q + 1][0], gf[2 * q + 1][1], gf[2 * q + 1][2], gf[2 * q + 1][3], l2, l3);
        }
        const int p = base + 8 * h + 2 * t;
        float a[2], dl[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float lgv = d1[i] + d1[i + 2] + bias, dav = d2[i] + d2[i + 2];    // logit, d attn of (l = g, pixel p + i)
          a[i] = p + i < n ? __fmul_rn(__expf(lgv - M), rS) : 0.f;
          dl[i] = __fmul_rn(a[i], dav - D);                                       // d logit
          db += dl[i];
        }
        // GEMM2: dW^T[c][l] += x[p][c] * dlogit[l][p]
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(dl[0], bh0, bl0); split_tf32(dl[1], bh1, bl1);
        const float* xp = xs + p * C + 2 * g;
#pragma unroll
        for (int j = 0; j < C / 16; ++j) {
          const float2 u0 = *reinterpret_cast<const float2*>(xp + 16 * j);
          const float2 u1 = *reinterpret_cast<const float2*>(xp + C + 16 * j);
          uint32_t ah[4], al[4];
          split_tf32(u0.x, ah[0], al[0]); split_tf32(u0.y, ah[1], al[1]);
          split_tf32(u1.x, ah[2], al[2]); split_tf32(u1.y, ah[3], al[3]);
          mma_tf32(dwt[j], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
          mma_tf32(dwt[j], al[0], al[1], al[2], al[3], bh0, bh1);
          mma_tf32(dwt[j], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
        }
        // transpose (l = g; pixels 2t, 2t+1) -> (pixel g; l = 2t, 2t+1) across the warp
        const float a00 = __shfl_sync(0xffffffffu, a[0], src0), a01 = __shfl_sync(0xffffffffu, a[1], src0);
        const float a10 = __shfl_sync(0xffffffffu, a[0], src0 + 4), a11 = __shfl_sync(0xffffffffu, a[1], src0 + 4);
        const float e00 = __shfl_sync(0xffffffffu, dl[0], src0), e01 = __shfl_sync(0xffffffffu, dl[1], src0);
        const float e10 = __shfl_sync(0xffffffffu, dl[0], src0 + 4), e11 = __shfl_sync(0xffffffffu, dl[1], src0 + 4);
        at[h][0] = odd ? a01 : a00;  at[h][1] = odd ? a11 : a10;
        dlt[h][0] = odd ? e01 : e00; dlt[h][1] = odd ? e11 : e10;
      }
      // GEMM3: dX[p][c] = sum_l attn[l][p] gtok[l][c] + dlogit[l][p] W[l][c]; rows g = group A pixel g, g+8 = group B pixel g
      uint32_t sh0[4], sl0[4], sh1[4], sl1[4];
      split_tf32(at[0][0], sh0[0], sl0[0]); split_tf32(at[1][0], sh0[1], sl0[1]);
      split_tf32(at[0][1], sh0[2], sl0[2]); split_tf32(at[1][1], sh0[3], sl0[3]);
      split_tf32(dlt[0][0], sh1[0], sl1[0]); split_tf32(dlt[1][0], sh1[1], sl1[1]);
      split_tf32(dlt[0][1], sh1[2], sl1[2]); split_tf32(dlt[1][1], sh1[3], sl1[3]);
      const bool live_a = base + g < n, live_b = base + 8 + g < n;
#pragma unroll
      for (int q = 0; q < C / 16; ++q) {
        float o[2][4];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * q + jj;
          const uint4 v0 = vs[(j * 2 + 0) * 32 + lane], v1 = vs[(j * 2 + 1) * 32 + lane];
          float d3[4] = {0.f, 0.f, 0.f, 0.f}, d4[4] = {0.f, 0.f, 0.f, 0.f};
          mma_tf32(d3, sh0[0], sh0[1], sh0[2], sh0[3], v0.x, v0.y);
          mma_tf32(d4, sh1[0], sh1[1], sh1[2], sh1[3], v1.x, v1.y);
          mma_tf32(d3, sl0[0], sl0[1], sl0[2], sl0[3], v0.x, v0.y);
          mma_tf32(d4, sl1[0], sl1[1], sl1[2], sl1[3], v1.x, v1.y);
          mma_tf32(d3, sh0[0], sh0[1], sh0[2], sh0[3], v0.z, v0.w);
          mma_tf32(d4, sh1[0], sh1[1], sh1[2], sh1[3], v1.z, v1.w);
          o[0][2 * jj] = d3[0] + d4[0]; o[0][2 * jj + 1] = d3[1] + d4[1];     // group A pixel g, channels 16q + 4t + 2jj, +1
          o[1][2 * jj] = d3[2] + d4[2]; o[1][2 * jj + 1] = d3[3] + d4[3];     // group B pixel g
        }
        if (live_a) *reinterpret_cast<float4*>(gxb + (int64_t)(base + g) * C + 16 * q) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
        if (live_b) *reinterpret_cast<float4*>(gxb + (int64_t)(base + 8 + g) * C + 16 * q) = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
      }
    }
    db += __shfl_xor_sync(0xffffffffu, db, 1);
    db += __shfl_xor_sync(0xffffffffu, db, 2);
    float* dwp = dww + warp * TOK_L * C;
#pragma unroll
    for (int j = 0; j < C / 16; ++j) {
      *reinterpret_cast<float2*>(dwp + (2 * t) * C + 16 * j + 2 * g) = make_float2(dwt[j][0], dwt[j][2]);
      *reinterpret_cast<float2*>(dwp + (2 * t + 1) * C + 16 * j + 2 * g) = make_float2(dwt[j][1], dwt[j][3]);
    }
    if (t == 0) dbw[warp * TOK_L + g] = db;
    __syncthreads();
    float* out = part + ((int64_t)bk * geo.nchunks + blockIdx.x) * (TOK_L * C + TOK_L);
    for (int i = threadIdx.x; i < TOK_L * C + TOK_L; i += NT) {
      float v = 0.f;
      if (i < TOK_L * C) {
#pragma unroll
        for (int w = 0; w < NW; ++w) v += dww[w * TOK_L * C + i];
      } else {
#pragma unroll
        for (int w = 0; w < NW; ++w) v += dbw[w * TOK_L + i - TOK_L * C];
      }
      out[i] = v;
    }
  }
}

bool tok_mma_supported(int C) { return C == 16 || C == 32; }
int tok_mma_chunk_px(int C) { (void)C; return 512; }               // 32 / 64 KB of x per CTA (1024 and 256 measured slower)

template <typename K> static void tokm_allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

void tok_fwd_mma_launch(const float* x, const TokWarpSrc* src, const float* wa, const float* ba, float* part,
                        const TokGeom& g, int B, cudaStream_t st) {
  const dim3 grid(g.nchunks, 4 * B);
  const int NW = TOKM_FWD_THREADS / 32;
  const size_t smem = ((size_t)g.chunk * g.C + (size_t)NW * (TOK_L * g.C + 2 * TOK_L)) * sizeof(float);
  const TokWarpSrc s = src ? *src : TokWarpSrc{};
#define SMOW_TOKM_FWD(CC, FU)                                    \
  tokm_allow_smem(tok_fwd_mma_kernel<CC, 512, FU>, smem);        \
  tok_fwd_mma_kernel<CC, 512, FU><<<grid, TOKM_FWD_THREADS, smem, st>>>(x, s, wa, ba, part, g)
  if (g.C == 16) { if (src) { SMOW_TOKM_FWD(16, true); } else { SMOW_TOKM_FWD(16, false); } }
  else           { if (src) { SMOW_TOKM_FWD(32, true); } else { SMOW_TOKM_FWD(32, false); } }
#undef SMOW_TOKM_FWD
}

void tok_bwd_mma_launch(const float* gtok, const float* x, const TokWarpSrc* src, const float* wa, const float* ba,
                        const float* tokens, const float* stats, float* gx, float* part, const TokGeom& g, int B,
                        cudaStream_t st) {
  const dim3 grid(g.nchunks, 4 * B);
  const int NW = TOKM_BWD_THREADS / 32;
  const size_t smem = ((size_t)g.chunk * g.C + (size_t)(g.C / 8) * 2 * 32 * 4 + (size_t)NW * (TOK_L * g.C + TOK_L)) * sizeof(float);
  const TokWarpSrc s = src ? *src : TokWarpSrc{};
#define SMOW_TOKM_BWD(CC, FU)                                    \
  tokm_allow_smem(tok_bwd_mma_kernel<CC, 512, FU>, smem);        \
  tok_bwd_mma_kernel<CC, 512, FU><<<grid, TOKM_BWD_THREADS, smem, st>>>(gtok, x, s, wa, ba, tokens, stats, gx, part, g)
  if (g.C == 16) { if (src) { SMOW_TOKM_BWD(16, true); } else { SMOW_TOKM_BWD(16, false); } }
  else           { if (src) { SMOW_TOKM_BWD(32, true); } else { SMOW_TOKM_BWD(32, false); } }
#undef SMOW_TOKM_BWD
}

}  // namespace smow
