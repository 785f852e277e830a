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
// Thread = (pixel, 4-channel vector); the q = C/4 lanes of a pixel are adjacent.  A lane keeps the 8 x 4 weights of ITS
// channels in registers for the whole kernel and the per-pixel dot products are finished with q-lane butterflies.
// Everything is summed in a fixed order (no float atomics): results are bit-reproducible.
#include "common.cuh"

namespace smow {

constexpr int TOK_L = 8;            // token_len of both reference models
constexpr int TOK_CHUNK = 2048;     // pixels per CTA

struct TokGeom { int q, qshift, C; int64_t hw; int nchunks; };

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, __fmul_rn(a.x, b.x))));
}
// sum over the q adjacent lanes of a pixel (q a power of two <= 32)
__device__ __forceinline__ float lanes_sum(float v, int q) {
  for (int d = q >> 1; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// sum / max over the lanes of a warp that hold the same channel vector (stride q)
__device__ __forceinline__ float pixels_sum(float v, int q) {
  for (int d = 16; d >= q; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float pixels_max(float v, int q) {
  for (int d = 16; d >= q; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// ---- forward, pass 1: chunk-local softmax statistics and un-normalised partial tokens ------------------------------
// grid (nchunks, 4*B).  part: [bk][chunk][ m[8] | s[8] | T[8][C] ]
__global__ void __launch_bounds__(256)
tok_fwd_chunk_kernel(const float* __restrict__ x, const float* __restrict__ wa, const float* __restrict__ ba,
                     float* __restrict__ part, TokGeom g) {
  __shared__ float red[8][TOK_L];
  __shared__ float mfin[TOK_L];
  extern __shared__ float accs[];                       // [8 warps][q][L][4] + [8 warps][L]
  const int C = g.C, q = g.q;
  const int v = threadIdx.x & (q - 1), pin = threadIdx.x >> g.qshift, ppi = 256 >> g.qshift;   // pixels per iteration
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * TOK_CHUNK;
  const int64_t p1 = p0 + TOK_CHUNK < g.hw ? p0 + TOK_CHUNK : g.hw;
  const float* xb = x + (int64_t)bk * g.hw * C + v * 4;
  float4 w[TOK_L];
  float b8[TOK_L];
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) { w[l] = ldg4(wa + l * C + v * 4); b8[l] = __ldg(ba + l); }
  // phase 1: the chunk's maximum logit per token
  float m[TOK_L];
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) m[l] = -INFINITY;
  for (int64_t pb = p0; pb < p1; pb += ppi) {           // every lane of a warp runs the same number of iterations
    const int64_t p = pb + pin;
    const bool live = p < p1;
    const float4 xv = live ? ldg4(xb + p * C) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < TOK_L; ++l) {
      const float lg = lanes_sum(dot4(w[l], xv), q) + b8[l];
      if (live) m[l] = fmaxf(m[l], lg);
    }
  }
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) {
    const float mm = pixels_max(m[l], q);
    if (lane == 0) red[warp][l] = mm;
  }
  __syncthreads();
  if (threadIdx.x < TOK_L) {
    float mm = red[0][threadIdx.x];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i][threadIdx.x]);
    mfin[threadIdx.x] = mm;
  }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) m[l] = mfin[l];
  // phase 2: s[l] = sum exp(logit - m), T[l][c] = sum exp(logit - m) * x[c]   (x comes back from L1 / L2)
  float s[TOK_L];
  float4 acc[TOK_L];
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) { s[l] = 0.f; acc[l] = make_float4(0.f, 0.f, 0.f, 0.f); }
  for (int64_t pb = p0; pb < p1; pb += ppi) {
    const int64_t p = pb + pin;
    const bool live = p < p1;
    const float4 xv = live ? ldg4(xb + p * C) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < TOK_L; ++l) {
      const float lg = lanes_sum(dot4(w[l], xv), q) + b8[l];
      const float e = live ? expf(lg - m[l]) : 0.f;
      s[l] += e;
      acc[l].x = fmaf(e, xv.x, acc[l].x); acc[l].y = fmaf(e, xv.y, acc[l].y);
      acc[l].z = fmaf(e, xv.z, acc[l].z); acc[l].w = fmaf(e, xv.w, acc[l].w);
    }
  }
  // block reduction in a fixed order: lanes of equal v inside the warp, then the 8 warps
  float* sacc = accs;                                   // [8][q][L][4]
  float* ssum = accs + 8 * q * TOK_L * 4;               // [8][L]
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) {
    const float a0 = pixels_sum(acc[l].x, q), a1 = pixels_sum(acc[l].y, q), a2 = pixels_sum(acc[l].z, q),
                a3 = pixels_sum(acc[l].w, q), ss = pixels_sum(s[l], q);
    if (lane < q) *reinterpret_cast<float4*>(sacc + ((warp * q + lane) * TOK_L + l) * 4) = make_float4(a0, a1, a2, a3);
    if (lane == 0) ssum[warp * TOK_L + l] = ss;
  }
  __syncthreads();
  float* out = part + ((int64_t)bk * g.nchunks + blockIdx.x) * (2 * TOK_L + TOK_L * C);
  for (int i = threadIdx.x; i < TOK_L * C; i += 256) {  // i = l*C + c
    const int l = i / C, c = i - l * C;
    float t = 0.f;
    for (int wp = 0; wp < 8; ++wp) t += sacc[((wp * q + (c >> 2)) * TOK_L + l) * 4 + (c & 3)];
    out[2 * TOK_L + i] = t;
  }
  if (threadIdx.x < TOK_L) {
    float t = 0.f;
    for (int wp = 0; wp < 8; ++wp) t += ssum[wp * TOK_L + threadIdx.x];
    out[threadIdx.x] = mfin[threadIdx.x];
    out[TOK_L + threadIdx.x] = t;
  }
}

// ---- forward, pass 2: combine the chunks of one (pair, frame) -------------------------------------------------------
// grid 4*B, threads = L*C rounded up to 32.  tokens: [bk][L][C]; stats: [bk][ M[8] | 1/S[8] ]
__global__ void tok_fwd_combine_kernel(const float* __restrict__ part, float* __restrict__ tokens,
                                       float* __restrict__ stats, TokGeom g) {
  const int C = g.C, bk = blockIdx.x;
  const int stride = 2 * TOK_L + TOK_L * C;
  const float* pb = part + (int64_t)bk * g.nchunks * stride;
  for (int i = threadIdx.x; i < TOK_L * C; i += blockDim.x) {
    const int l = i / C;
    float M = -INFINITY;
    for (int k = 0; k < g.nchunks; ++k) M = fmaxf(M, pb[k * stride + l]);
    float S = 0.f, T = 0.f;
    for (int k = 0; k < g.nchunks; ++k) {
      const float sc = expf(pb[k * stride + l] - M);
      S = fmaf(pb[k * stride + TOK_L + l], sc, S);
      T = fmaf(pb[k * stride + 2 * TOK_L + i], sc, T);
    }
    tokens[(int64_t)bk * TOK_L * C + i] = __fdiv_rn(T, S);
    if (i == l * C) { stats[bk * 2 * TOK_L + l] = M; stats[bk * 2 * TOK_L + TOK_L + l] = __fdiv_rn(1.f, S); }
  }
}

// ---- backward: one pass ---------------------------------------------------------------------------------------------
// grid (nchunks, 4*B).  part: [bk][chunk][ dW[8][C] | db[8] ]
__global__ void __launch_bounds__(256)
tok_bwd_chunk_kernel(const float* __restrict__ gtok, const float* __restrict__ x, const float* __restrict__ wa,
                     const float* __restrict__ ba, const float* __restrict__ tokens, const float* __restrict__ stats,
                     float* __restrict__ gx, float* __restrict__ part, TokGeom g) {
  __shared__ float dsum[TOK_L];
  extern __shared__ float accs[];                       // [8 warps][q][L][4] + [8 warps][L]
  const int C = g.C, q = g.q;
  const int v = threadIdx.x & (q - 1), pin = threadIdx.x >> g.qshift, ppi = 256 >> g.qshift;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * TOK_CHUNK;
  const int64_t p1 = p0 + TOK_CHUNK < g.hw ? p0 + TOK_CHUNK : g.hw;
  const float* xb = x + (int64_t)bk * g.hw * C + v * 4;
  float* gxb = gx + (int64_t)bk * g.hw * C + v * 4;
  const float* gt = gtok + (int64_t)bk * TOK_L * C;
  // D[l] = sum_p attn[l,p] * dattn[l,p] = <gtok[l,:], tokens[l,:]>
  if (threadIdx.x < TOK_L) {
    const float* tk = tokens + (int64_t)bk * TOK_L * C + threadIdx.x * C;
    float d = 0.f;
    for (int c = 0; c < C; ++c) d = fmaf(__ldg(gt + threadIdx.x * C + c), __ldg(tk + c), d);
    dsum[threadIdx.x] = d;
  }
  __syncthreads();
  float4 w[TOK_L], gv[TOK_L], dw[TOK_L];
  float b8[TOK_L], M[TOK_L], rS[TOK_L], D[TOK_L], db[TOK_L];
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) {
    w[l] = ldg4(wa + l * C + v * 4);
    gv[l] = ldg4(gt + l * C + v * 4);
    b8[l] = __ldg(ba + l);
    M[l] = __ldg(stats + bk * 2 * TOK_L + l);
    rS[l] = __ldg(stats + bk * 2 * TOK_L + TOK_L + l);
    D[l] = dsum[l];
    dw[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[l] = 0.f;
  }
  for (int64_t pb = p0; pb < p1; pb += ppi) {
    const int64_t p = pb + pin;
    const bool live = p < p1;
    const float4 xv = live ? ldg4(xb + p * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < TOK_L; ++l) {
      const float lg = lanes_sum(dot4(w[l], xv), q) + b8[l];
      const float da = lanes_sum(dot4(gv[l], xv), q);             // d attn[l,p]
      const float a = live ? __fmul_rn(expf(lg - M[l]), rS[l]) : 0.f;
      const float dl = __fmul_rn(a, da - D[l]);                   // d logit[l,p]
      o.x = fmaf(a, gv[l].x, fmaf(dl, w[l].x, o.x)); o.y = fmaf(a, gv[l].y, fmaf(dl, w[l].y, o.y));
      o.z = fmaf(a, gv[l].z, fmaf(dl, w[l].z, o.z)); o.w = fmaf(a, gv[l].w, fmaf(dl, w[l].w, o.w));
      dw[l].x = fmaf(dl, xv.x, dw[l].x); dw[l].y = fmaf(dl, xv.y, dw[l].y);
      dw[l].z = fmaf(dl, xv.z, dw[l].z); dw[l].w = fmaf(dl, xv.w, dw[l].w);
      db[l] += dl;
    }
    if (live) *reinterpret_cast<float4*>(gxb + p * C) = o;
  }
  float* sacc = accs;
  float* ssum = accs + 8 * q * TOK_L * 4;
#pragma unroll
  for (int l = 0; l < TOK_L; ++l) {
    const float a0 = pixels_sum(dw[l].x, q), a1 = pixels_sum(dw[l].y, q), a2 = pixels_sum(dw[l].z, q),
                a3 = pixels_sum(dw[l].w, q), ss = pixels_sum(db[l], q);
    if (lane < q) *reinterpret_cast<float4*>(sacc + ((warp * q + lane) * TOK_L + l) * 4) = make_float4(a0, a1, a2, a3);
    if (lane == 0) ssum[warp * TOK_L + l] = ss;
  }
  __syncthreads();
  float* out = part + ((int64_t)bk * g.nchunks + blockIdx.x) * (TOK_L * C + TOK_L);
  for (int i = threadIdx.x; i < TOK_L * C; i += 256) {
    const int l = i / C, c = i - l * C;
    float t = 0.f;
    for (int wp = 0; wp < 8; ++wp) t += sacc[((wp * q + (c >> 2)) * TOK_L + l) * 4 + (c & 3)];
    out[i] = t;
  }
  if (threadIdx.x < TOK_L) {
    float t = 0.f;
    for (int wp = 0; wp < 8; ++wp) t += ssum[wp * TOK_L + threadIdx.x];
    out[TOK_L * C + threadIdx.x] = t;
  }
}

// gwa[L][C], gba[L] = sum over every (pair, frame, chunk) partial, in index order.  grid = ceil((L*C + L) / 128)
__global__ void tok_bwd_combine_kernel(const float* __restrict__ part, float* __restrict__ gwa, float* __restrict__ gba,
                                       int n_part, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, n = TOK_L * C + TOK_L;
  if (i >= n) return;
  float t = 0.f;
  for (int k = 0; k < n_part; ++k) t += part[(int64_t)k * n + i];
  if (i < TOK_L * C) gwa[i] = t; else gba[i - TOK_L * C] = t;
}

static int tok_geom(TokGeom& g, int B, int C, int64_t hw, const void* x, int dtype, int layout) {
  if (B <= 0 || C <= 0 || hw <= 0 || !x) return fail(SMOW_EINVAL, "tokenizer: bad shape / null pointer");
  if (dtype != SMOW_F32 || layout != SMOW_NDHWC)
    return fail(SMOW_EDTYPE, "tokenizer: built for fp32 channels_last_3d (NDHWC) stacks only");
  g.q = C / 4; g.qshift = -1; g.C = C; g.hw = hw;
  for (int s = 0; s < 6; ++s) if ((1 << s) == g.q) g.qshift = s;
  if (C % 4 || g.qshift < 0) return fail(SMOW_EINVAL, "tokenizer: C/4 must be a power of two <= 32 (got C = %d)", C);
  if (!aligned16(x)) return fail(SMOW_EALIGN, "tokenizer: stack not 16 B aligned");
  g.nchunks = (int)((hw + TOK_CHUNK - 1) / TOK_CHUNK);
  if ((int64_t)4 * B > 65535) return fail(SMOW_ERANGE, "tokenizer: batch too large for one launch");
  return 0;
}
static size_t tok_smem(const TokGeom& g) { return (size_t)(8 * g.q * TOK_L * 4 + 8 * TOK_L) * sizeof(float); }

}  // namespace smow

using namespace smow;

extern "C" {

int64_t smow_tokenizer_workspace_bytes(int B, int C, int64_t hw) {
  const int64_t nchunks = (hw + TOK_CHUNK - 1) / TOK_CHUNK;
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
  tok_fwd_chunk_kernel<<<dim3(g.nchunks, 4 * B), 256, tok_smem(g), st>>>((const float*)x, wa, ba, part, g);
  int th = TOK_L * C; th = th > 256 ? 256 : ((th + 31) / 32) * 32;
  tok_fwd_combine_kernel<<<4 * B, th, 0, st>>>(part, tokens, stats, g);
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
  tok_bwd_chunk_kernel<<<dim3(g.nchunks, 4 * B), 256, tok_smem(g), st>>>(gtokens, (const float*)x, wa, ba, tokens, stats,
                                                                         (float*)gx, part, g);
  const int n = TOK_L * C + TOK_L;
  tok_bwd_combine_kernel<<<(n + 127) / 128, 128, 0, st>>>(part, gwa, gba, 4 * B * g.nchunks, C);
  count_launch(2);
  return check_launch("tokenizer_bwd");
}

}  // extern "C"
