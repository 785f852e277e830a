// Row N4: the decoder's cyclic temporal frame mix (reference models/SMOW_Net.py:121-139,
// models/SMOW_Net_LW.py:121-137,160-175) for the channel counts of the large decoder levels (C = 16, 28, 32, 64):
//     out[b, f, p, :] = in[b, f, p, :] @ M0  +  in[b, (f + shift) % 4, p, :] @ M1[(f + own_off) % 4]
// forward:  in = X,  out = Y,  M0 = W_shared, M1 = W_own,   shift = 1, own_off = 1   (Y_j = X_j W5 + X_{j+1} W_{j+1})
// d(input): in = gY, out = gX, M0 = W_shared^T, M1 = W_own^T, shift = 3, own_off = 0 (gX_k = gY_k W5^T + gY_{k-1} W_k^T)
// The reference composes it from 4 slices, ten 1x1x1 convolutions, 4 adds and a concat; cuBLAS needs 5 GEMM passes
// that re-read X and read-modify-write Y.  Here every input frame tile is read once per use and Y is written once.
// Weight gradients: per CTA D0 = X_f^T G_f (-> dW_shared) and D1 = X_f^T G_{f-1} (-> dW_own[f]) over a pixel range,
// summed over CTAs in index order by a second kernel (deterministic).  fp32, NDHWC, plain FFMA (K = C is far too small
// for tensor cores to matter: the op is HBM bound, ~2C flops per byte).
#include "common.cuh"

namespace smow {

constexpr int MIX_PX = 64;                  // pixels per tile

// ---- apply (forward and d(input)) ----------------------------------------------------------------------------------
// grid (ceil(hw / 64), 4*B); block 16 * C/4 threads: thread = (4 pixels, 4 output channels)
template <int C>
__global__ void __launch_bounds__(4 * C)
mix_apply_kernel(const float* __restrict__ in, const float* __restrict__ m0, const float* __restrict__ m1,
                 float* __restrict__ out, int64_t hw, int shift, int own_off) {
  constexpr int Q = C / 4, NT = 16 * Q, LD = MIX_PX + 4;
  extern __shared__ __align__(16) float sm[];
  float* xa = sm;                           // [C][LD]  frame f, transposed (channel-major)
  float* xb = xa + C * LD;                  // [C][LD]  frame (f + shift) % 4
  float* wa = xb + C * LD;                  // [C][C]   M0, row = input channel
  float* wb = wa + C * C;                   // [C][C]   M1[(f + own_off) % 4]
  const int bf = blockIdx.y, b = bf >> 2, f = bf & 3;
  const int fn = (f + shift) & 3, g = (f + own_off) & 3;
  const int64_t p0 = (int64_t)blockIdx.x * MIX_PX;
  const float* ia = in + ((int64_t)(b * 4 + f) * hw + p0) * C;
  const float* ib = in + ((int64_t)(b * 4 + fn) * hw + p0) * C;
  const int npx = hw - p0 < MIX_PX ? (int)(hw - p0) : MIX_PX;
  for (int i = threadIdx.x; i < MIX_PX * Q; i += NT) {
    const int px = i / Q, v = i - px * Q;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
    if (px < npx) {
      a = __ldg(reinterpret_cast<const float4*>(ia + (int64_t)px * C + 4 * v));
      c = __ldg(reinterpret_cast<const float4*>(ib + (int64_t)px * C + 4 * v));
    }
    xa[(4 * v + 0) * LD + px] = a.x; xa[(4 * v + 1) * LD + px] = a.y;
    xa[(4 * v + 2) * LD + px] = a.z; xa[(4 * v + 3) * LD + px] = a.w;
    xb[(4 * v + 0) * LD + px] = c.x; xb[(4 * v + 1) * LD + px] = c.y;
    xb[(4 * v + 2) * LD + px] = c.z; xb[(4 * v + 3) * LD + px] = c.w;
  }
  const float* m1g = m1 + (int64_t)g * C * C;
  for (int i = threadIdx.x; i < C * Q; i += NT) {
    reinterpret_cast<float4*>(wa)[i] = __ldg(reinterpret_cast<const float4*>(m0) + i);
    reinterpret_cast<float4*>(wb)[i] = __ldg(reinterpret_cast<const float4*>(m1g) + i);
  }
  __syncthreads();
  const int pg = threadIdx.x / Q, cv = threadIdx.x - pg * Q;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int ci = 0; ci < C; ++ci) {
    const float4 a = *reinterpret_cast<const float4*>(xa + ci * LD + 4 * pg);
    const float4 c = *reinterpret_cast<const float4*>(xb + ci * LD + 4 * pg);
    const float4 w0 = *reinterpret_cast<const float4*>(wa + ci * C + 4 * cv);
    const float4 w1 = *reinterpret_cast<const float4*>(wb + ci * C + 4 * cv);
    const float av[4] = {a.x, a.y, a.z, a.w}, cvv[4] = {c.x, c.y, c.z, c.w};
    const float u[4] = {w0.x, w0.y, w0.z, w0.w}, t[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(cvv[i], t[j], fmaf(av[i], u[j], acc[i][j]));
  }
  float* o = out + ((int64_t)(b * 4 + f) * hw + p0) * C + 4 * cv;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int px = 4 * pg + i;
    if (px < npx) *reinterpret_cast<float4*>(o + (int64_t)px * C) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// ---- weight gradients ----------------------------------------------------------------------------------------------
// grid (nranges, 4*B); block 256.  part: [bf][range][2][C][C]  (D0 = X_f^T G_f, D1 = X_f^T G_{(f+3)%4})
template <int C>
__global__ void __launch_bounds__(256)
mix_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ part, int64_t hw,
                 int range_px, int nranges) {
  constexpr int Q = C / 4, T2 = Q * Q, S = 256 / T2 > 0 ? 256 / T2 : 1;
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                           // [64][C] row-major tiles
  float* gs = xs + MIX_PX * C;
  float* gp = gs + MIX_PX * C;
  const int bf = blockIdx.y, b = bf >> 2, f = bf & 3, fp = (f + 3) & 3;
  const int64_t r0 = (int64_t)blockIdx.x * range_px;
  const int64_t r1 = r0 + range_px < hw ? r0 + range_px : hw;
  const float* xb = x + (int64_t)(b * 4 + f) * hw * C;
  const float* gb = gy + (int64_t)(b * 4 + f) * hw * C;
  const float* hb = gy + (int64_t)(b * 4 + fp) * hw * C;
  const int tile = threadIdx.x % T2, s = threadIdx.x / T2;       // s >= S: idle thread (C = 28: 245 of 256 work)
  const int ci4 = tile / Q, co4 = tile - ci4 * Q;
  float a0[4][4], a1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a0[i][j] = a1[i][j] = 0.f;
  for (int64_t p0 = r0; p0 < r1; p0 += MIX_PX) {
    const int npx = r1 - p0 < MIX_PX ? (int)(r1 - p0) : MIX_PX;
    __syncthreads();
    for (int i = threadIdx.x; i < MIX_PX * Q; i += 256) {
      const int px = i / Q;
      float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vg = vx, vh = vx;      // pixels past the range add exact zeros
      if (px < npx) {
        const int64_t e = (p0 + px) * C + 4 * (i - px * Q);
        vx = __ldg(reinterpret_cast<const float4*>(xb + e));
        vg = __ldg(reinterpret_cast<const float4*>(gb + e));
        vh = __ldg(reinterpret_cast<const float4*>(hb + e));
      }
      reinterpret_cast<float4*>(xs)[i] = vx;
      reinterpret_cast<float4*>(gs)[i] = vg;
      reinterpret_cast<float4*>(gp)[i] = vh;
    }
    __syncthreads();
    if (s < S) {
      for (int px = s; px < MIX_PX; px += S) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + px * C + 4 * ci4);
        const float4 gv = *reinterpret_cast<const float4*>(gs + px * C + 4 * co4);
        const float4 hv = *reinterpret_cast<const float4*>(gp + px * C + 4 * co4);
        const float xr[4] = {xv.x, xv.y, xv.z, xv.w}, gr[4] = {gv.x, gv.y, gv.z, gv.w}, hr[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) { a0[i][j] = fmaf(xr[i], gr[j], a0[i][j]); a1[i][j] = fmaf(xr[i], hr[j], a1[i][j]); }
      }
    }
  }
  // reduce the S pixel-splits in a fixed order through shared memory, then one partial per CTA
  __syncthreads();
  float* red = sm;                          // [S][T2][32]  (the launch sizes shared memory for it)
  if (s < S) {
    float* r = red + ((int64_t)s * T2 + tile) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { r[i * 4 + j] = a0[i][j]; r[16 + i * 4 + j] = a1[i][j]; }
  }
  __syncthreads();
  float* outp = part + ((int64_t)bf * nranges + blockIdx.x) * 2 * C * C;
  for (int e = threadIdx.x; e < 2 * C * C; e += 256) {
    const int m = e / (C * C), rem = e - m * C * C, ci = rem / C, co = rem - ci * C;
    const int tl = (ci >> 2) * Q + (co >> 2), slot = m * 16 + (ci & 3) * 4 + (co & 3);
    float t = 0.f;
    for (int k = 0; k < S; ++k) t += red[((int64_t)k * T2 + tl) * 32 + slot];
    outp[e] = t;
  }
}

// gw: [5][C][C] = dW_shared, dW_own[0..3].  grid (ceil(C*C / 32), 5); block 256 = 32 elements x 8 partial lanes
__global__ void __launch_bounds__(256)
mix_wgrad_combine_kernel(const float* __restrict__ part, float* __restrict__ gw, int CC, int n_bf, int nranges) {
  __shared__ float red[8][32];
  const int m = blockIdx.y, el = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + el;
  float t = 0.f;
  if (e < CC) {
    const int P = n_bf * nranges;
    for (int p = pl; p < P; p += 8) {
      const int f = (p / nranges) & 3;
      if (m == 0) t += part[(int64_t)p * 2 * CC + e];
      else if (f == m - 1) t += part[(int64_t)p * 2 * CC + CC + e];
    }
  }
  red[pl][el] = t;
  __syncthreads();
  if (pl == 0 && e < CC) {
    float r = 0.f;
    for (int k = 0; k < 8; ++k) r += red[k][el];
    gw[(int64_t)m * CC + e] = r;
  }
}

static int mix_range_px(int64_t hw) {
  int64_t r = hw / 8;
  if (r < 256) r = 256;
  r = (r + MIX_PX - 1) / MIX_PX * MIX_PX;
  return (int)r;
}
static bool mix_supported(int C) { return C == 16 || C == 28 || C == 32 || C == 64; }

template <typename K> static void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int C>
static void launch_apply(const float* in, const float* m0, const float* m1, float* out, int B, int64_t hw, int shift,
                         int own_off, cudaStream_t st) {
  const size_t smem = (size_t)(2 * C * (MIX_PX + 4) + 2 * C * C) * sizeof(float);
  allow_smem(mix_apply_kernel<C>, smem);
  mix_apply_kernel<C><<<dim3((unsigned)((hw + MIX_PX - 1) / MIX_PX), 4 * B), 4 * C, smem, st>>>(in, m0, m1, out, hw, shift, own_off);
}
template <int C>
static void launch_wgrad(const float* x, const float* gy, float* part, int B, int64_t hw, int range_px, int nranges,
                         cudaStream_t st) {
  // three [64][C] tiles, re-used at the end as the [S][T2][32] reduction scratch (256 * 32 floats at most)
  const size_t tiles = (size_t)3 * MIX_PX * C, scratch = 256 * 32;
  const size_t smem = (tiles > scratch ? tiles : scratch) * sizeof(float);
  allow_smem(mix_wgrad_kernel<C>, smem);
  mix_wgrad_kernel<C><<<dim3(nranges, 4 * B), 256, smem, st>>>(x, gy, part, hw, range_px, nranges);
}

}  // namespace smow

using namespace smow;

extern "C" {

int smow_frame_mix_supported(int C) { return mix_supported(C) ? 1 : 0; }

int smow_frame_mix_apply(const float* in, const float* m0, const float* m1, float* out, int B, int C, int64_t hw,
                         int shift, int own_off, void* stream) {
  if (!in || !m0 || !m1 || !out || B <= 0 || hw <= 0) return fail(SMOW_EINVAL, "frame_mix: bad shape / null pointer");
  if (!mix_supported(C)) return fail(SMOW_EDTYPE, "frame_mix: built for C in {16, 28, 32, 64} (got %d)", C);
  if (!aligned16(in) || !aligned16(out) || !aligned16(m0) || !aligned16(m1)) return fail(SMOW_EALIGN, "frame_mix: 16 B alignment");
  if ((int64_t)4 * B > 65535) return fail(SMOW_ERANGE, "frame_mix: batch too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  shift &= 3; own_off &= 3;
  switch (C) {
    case 16: launch_apply<16>(in, m0, m1, out, B, hw, shift, own_off, st); break;
    case 28: launch_apply<28>(in, m0, m1, out, B, hw, shift, own_off, st); break;
    case 32: launch_apply<32>(in, m0, m1, out, B, hw, shift, own_off, st); break;
    default: launch_apply<64>(in, m0, m1, out, B, hw, shift, own_off, st); break;
  }
  count_launch();
  return check_launch("frame_mix_apply");
}

int64_t smow_frame_mix_wgrad_workspace_bytes(int B, int C, int64_t hw) {
  if (B <= 0 || C <= 0 || hw <= 0) return 0;
  const int range_px = mix_range_px(hw);
  const int64_t nranges = (hw + range_px - 1) / range_px;
  return (int64_t)4 * B * nranges * 2 * C * C * (int64_t)sizeof(float);
}

int smow_frame_mix_wgrad(const float* x, const float* gy, float* gw, int B, int C, int64_t hw, void* ws,
                         int64_t ws_bytes, void* stream) {
  if (!x || !gy || !gw || B <= 0 || hw <= 0) return fail(SMOW_EINVAL, "frame_mix: bad shape / null pointer");
  if (!mix_supported(C)) return fail(SMOW_EDTYPE, "frame_mix: built for C in {16, 28, 32, 64} (got %d)", C);
  if (!aligned16(x) || !aligned16(gy) || !ws || !aligned16(ws) || ws_bytes < smow_frame_mix_wgrad_workspace_bytes(B, C, hw))
    return fail(SMOW_EINVAL, "frame_mix: workspace of smow_frame_mix_wgrad_workspace_bytes() bytes required");
  if ((int64_t)4 * B > 65535) return fail(SMOW_ERANGE, "frame_mix: batch too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  const int range_px = mix_range_px(hw);
  const int nranges = (int)((hw + range_px - 1) / range_px);
  float* part = reinterpret_cast<float*>(ws);
  switch (C) {
    case 16: launch_wgrad<16>(x, gy, part, B, hw, range_px, nranges, st); break;
    case 28: launch_wgrad<28>(x, gy, part, B, hw, range_px, nranges, st); break;
    case 32: launch_wgrad<32>(x, gy, part, B, hw, range_px, nranges, st); break;
    default: launch_wgrad<64>(x, gy, part, B, hw, range_px, nranges, st); break;
  }
  mix_wgrad_combine_kernel<<<dim3((C * C + 31) / 32, 5), 256, 0, st>>>(part, gw, C * C, 4 * B, nranges);
  count_launch(2);
  return check_launch("frame_mix_wgrad");
}

}  // extern "C"
