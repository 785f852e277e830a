// BatchNorm of the decoder blocks, folded into the kernels around it (rows A3 + A4 + N4; reference
// models/SMOW_Net.py:135-137 `conv_trans_block_3d.forward`: frame mix -> self.batch -> self.leaky, and
// models/SMOW_Net_LW.py:160-175 `conv_block_2_3d.forward`):
//   * the batch statistics (sum y, sum y^2 per channel) are by-products of the tcgen05 frame-mix epilogue
//     (frame_mix_tc.cu), one partial row per CTA;
//   * bn_finalize_kernel turns them into scale = gamma * invstd, shift = beta - mean * scale, updates running_mean /
//     running_var exactly like nn.BatchNorm3d in training mode (momentum, unbiased running variance);
//   * the normalisation itself is applied by the launch that writes the concat buffer (tlerp_cat.cu: affine + LeakyReLU);
//   * backward: one reduction pass over (gradient slice, y) for sum(du) and sum(du * xhat), a finalize, and the apply pass
//     inside the fused lerp + concat backward.  cuDNN's two BatchNorm passes per direction and the stand-alone LeakyReLU
//     kernels are gone.  fp32; statistics are reduced in double precision; fixed summation order (deterministic).
#include "common.cuh"

namespace smow {

// Sum of the per-CTA partial rows of one channel: one CTA (256 threads) per channel, every thread owns a strided subset of the
// partials (2-3 independent loads instead of a chain of ~20 L2 round trips per lane), fixed shuffle tree, then the 8 warp sums
// in index order: deterministic, double precision.
__device__ __forceinline__ bool channel_sums(const float* __restrict__ parts, int nparts, int C, int c, double& s1, double& s2) {
  __shared__ double red[2][8];
  s1 = 0.0; s2 = 0.0;
  for (int k = threadIdx.x; k < nparts; k += 256) {
    s1 += (double)parts[((size_t)k * 2 + 0) * C + c];
    s2 += (double)parts[((size_t)k * 2 + 1) * C + c];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x != 0) return false;
  s1 = 0.0; s2 = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { s1 += red[0][w]; s2 += red[1][w]; }
  return true;
}

// bn: [6][C] = scale | shift | mean | invstd | k1 | k2 (k1, k2 are written by the backward)
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float* __restrict__ parts, int nparts, int C, double count, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                   float momentum, float eps, float* __restrict__ bn) {
  const int c = blockIdx.x;
  double s1, s2;
  if (!channel_sums(parts, nparts, C, c, s1, s2)) return;
  const double mean = s1 / count;
  double var = s2 / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float scale = g * invstd;
  bn[c] = scale;
  bn[C + c] = b - (float)mean * scale;
  bn[2 * C + c] = (float)mean;
  bn[3 * C + c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// partial sums of du and du * xhat over the rows of the decoder half.  part: [block][2][Cd]
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ gcat, const float* __restrict__ y, const float* __restrict__ bn,
                         int64_t rows, int Cd, int Ct, float slope, float* __restrict__ part) {
  extern __shared__ __align__(16) float sm[];            // [rows_per_iter][q][8]
  const int q = Cd >> 2, rpi = 256 / q;
  const int rl = threadIdx.x / q, v = threadIdx.x - rl * q;
  float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
  if (rl < rpi) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(bn) + v), sh = __ldg(reinterpret_cast<const float4*>(bn + Cd) + v);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(bn + 2 * Cd) + v), is = __ldg(reinterpret_cast<const float4*>(bn + 3 * Cd) + v);
    const int64_t step = (int64_t)gridDim.x * rpi;
    int64_t r = (int64_t)blockIdx.x * rpi + rl;
#define SMOW_BN_ACC(gv, yv, f)                                                   \
  {                                                                              \
    const float u = fmaf(yv.f, sc.f, sh.f), du = u > 0.f ? gv.f : gv.f * slope;  \
    a1.f += du;                                                                  \
    a2.f = fmaf(du, (yv.f - mu.f) * is.f, a2.f);                                 \
  }
#define SMOW_BN_ROW(gv, yv) SMOW_BN_ACC(gv, yv, x) SMOW_BN_ACC(gv, yv, y) SMOW_BN_ACC(gv, yv, z) SMOW_BN_ACC(gv, yv, w)
    for (; r + 3 * step < rows; r += 4 * step) {                    // four rows in flight per thread
      float4 g[4], yy[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        g[k] = __ldg(reinterpret_cast<const float4*>(gcat + (r + k * step) * Ct) + v);
        yy[k] = __ldg(reinterpret_cast<const float4*>(y + (r + k * step) * Cd) + v);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { SMOW_BN_ROW(g[k], yy[k]) }
    }
    for (; r < rows; r += step) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gcat + r * Ct) + v);
      const float4 yy = __ldg(reinterpret_cast<const float4*>(y + r * Cd) + v);
      SMOW_BN_ROW(g, yy)
    }
#undef SMOW_BN_ROW
#undef SMOW_BN_ACC
    float* dst = sm + ((size_t)rl * q + v) * 8;
    *reinterpret_cast<float4*>(dst) = a1;
    *reinterpret_cast<float4*>(dst + 4) = a2;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * Cd; e += 256) {
    const int which = e / Cd, c = e - which * Cd;
    float t = 0.f;
    for (int k = 0; k < rpi; ++k) t += sm[((size_t)k * q + (c >> 2)) * 8 + which * 4 + (c & 3)];
    part[((size_t)blockIdx.x * 2 + which) * Cd + c] = t;
  }
}

__global__ void __launch_bounds__(256)
bn_act_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, double count, float* __restrict__ bn,
                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x;
  double s1, s2;
  if (!channel_sums(part, nparts, C, c, s1, s2)) return;
  bn[4 * C + c] = (float)(s1 / count);
  bn[5 * C + c] = (float)(s2 / count);
  if (dbeta) dbeta[c] = (float)s1;
  if (dgamma) dgamma[c] = (float)s2;
}

static int bn_reduce_blocks(int64_t rows, int Cd) {
  const int rpi = 256 / (Cd / 4);
  const int64_t want = (rows + rpi - 1) / rpi;
  const int cap = device_info().sms * 4;
  return (int)(want < cap ? want : cap);
}

}  // namespace smow

using namespace smow;

extern "C" {

int smow_bn_finalize(const float* parts, int nparts, int C, int64_t count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, float* bn, void* stream) {
  if (!parts || !bn || nparts <= 0 || C <= 0 || count <= 0) return fail(SMOW_EINVAL, "bn_finalize: bad argument");
  bn_finalize_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(parts, nparts, C, (double)count, gamma, beta,
                                                                        running_mean, running_var, momentum, eps, bn);
  count_launch();
  return check_launch("bn_finalize");
}

int64_t smow_bn_act_bwd_workspace_bytes(int B, int Cd, int64_t hw) {
  if (B <= 0 || Cd <= 0 || Cd % 4 || Cd > 1024 || hw <= 0) return 0;
  return (int64_t)bn_reduce_blocks((int64_t)B * 4 * hw, Cd) * 2 * Cd * (int64_t)sizeof(float);
}

// first half of the BatchNorm + LeakyReLU backward: fills bn[4] / bn[5] (k1, k2), dgamma, dbeta
int smow_bn_act_bwd_reduce(const float* gcat, const float* y, float* bn, float* dgamma, float* dbeta, int B, int Cd, int Cs,
                           int64_t hw, float slope, void* ws, int64_t ws_bytes, void* stream) {
  if (!gcat || !y || !bn || B <= 0 || Cd <= 0 || Cd % 4 || Cs % 4 || Cd > 1024 || hw <= 0)
    return fail(SMOW_EINVAL, "bn_act_bwd: bad argument (Cd, Cs multiples of 4, Cd <= 1024)");
  if (!aligned16(gcat) || !aligned16(y) || !aligned16(bn) || !ws || !aligned16(ws) ||
      ws_bytes < smow_bn_act_bwd_workspace_bytes(B, Cd, hw))
    return fail(SMOW_EALIGN, "bn_act_bwd: 16 B alignment / workspace of smow_bn_act_bwd_workspace_bytes() bytes");
  const int64_t rows = (int64_t)B * 4 * hw;
  const int nb = bn_reduce_blocks(rows, Cd);
  const int q = Cd / 4, rpi = 256 / q;
  const size_t smem = (size_t)rpi * q * 8 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  float* part = reinterpret_cast<float*>(ws);
  bn_act_bwd_reduce_kernel<<<nb, 256, smem, st>>>(gcat, y, bn, rows, Cd, Cd + Cs, slope, part);
  bn_act_bwd_finalize_kernel<<<Cd, 256, 0, st>>>(part, nb, Cd, (double)rows, bn, dgamma, dbeta);
  count_launch(2);
  return check_launch("bn_act_bwd_reduce");
}

}  // extern "C"
