// A3 + A4: temporal 2 -> 4 linear upsample of an encoder skip, written straight into
// the decoder's channel concat, and its backward (sm_100a, pure HBM streaming).
//
// Reference (models/SMOW_Net.py:64-73,78-94): F.interpolate(x, size=(4,h,w), 'trilinear',
// align_corners=True) then torch.cat([dec, x_up], dim=1).  With unchanged h,w the 8-tap
// trilinear kernel degenerates to out[t] = (1-l_t) T1 + l_t T2, l = {0, 1/3, 2/3, 1} in
// fp32 (ATen UpSampleTrilinear3d: rdepth = (2-1)/(4-1); frames 0 and 3 are exact copies).
#include <type_traits>
#include "common.cuh"

namespace smow {

template <typename T> struct Vec;   // 16-byte vector of T
template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __nv_bfloat16 v[8];
};

template <typename T> __device__ __forceinline__ Vec<T> ldv(const T* p) {
  Vec<T> r;
  *reinterpret_cast<uint4*>(r.v) = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
template <typename T> __device__ __forceinline__ Vec<T> ldv_stream(const T* p) {
  Vec<T> r;
  *reinterpret_cast<uint4*>(r.v) = __ldcs(reinterpret_cast<const uint4*>(p));
  return r;
}
template <typename T> __device__ __forceinline__ void stv(T* p, const Vec<T>& r) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
}

// (row, element) of flat vector index i; 32-bit division when the problem allows it (always, in practice)
__device__ __forceinline__ void split_index(int64_t i, int64_t per_row, bool small, int64_t& row, int64_t& col) {
  if (small) {
    const uint32_t r = (uint32_t)i / (uint32_t)per_row;
    row = r; col = (uint32_t)i - r * (uint32_t)per_row;
  } else {
    row = i / per_row; col = i - row * per_row;
  }
}

// exact division of a dividend < 2^31 by a runtime constant: q = (n * m) >> sh, m = floor(2^sh / d) + 1,
// sh = 31 + ceil(log2 d)  (error term n*e / (d * 2^sh) < 1/d).  Three instructions instead of ~20.
struct FastDiv { uint32_t d, m; int sh; };
static FastDiv make_fastdiv(int64_t d64) {
  FastDiv f;
  f.d = (uint32_t)d64;
  int L = 0;
  while ((1ull << L) < (uint64_t)d64) ++L;
  f.sh = 31 + L;
  f.m = (uint32_t)(((1ull << f.sh) / (uint64_t)d64) + 1ull);
  return f;
}
__device__ __forceinline__ void fast_split(uint32_t i, const FastDiv& f, uint32_t& row, uint32_t& col) {
  row = (uint32_t)(((uint64_t)i * f.m) >> f.sh);
  col = i - row * f.d;
}

struct LerpW { float a1, b1, a2, b2; };
__device__ __forceinline__ LerpW lerp_weights() {
  LerpW w;
  w.b1 = SMOW_LAMBDA1; w.a1 = __fsub_rn(1.f, SMOW_LAMBDA1);
  w.b2 = SMOW_LAMBDA2; w.a2 = __fsub_rn(1.f, SMOW_LAMBDA2);
  return w;
}

// One launch: blocks [0, lerp_blocks) lerp the skip into cat[:, Cd:], the rest copy dec
// into cat[:, :Cd].  VEC = elements per thread-access (16 B vectors, or 1 for odd shapes).
template <typename T, bool VECTOR>
__global__ void __launch_bounds__(256)
tlerp_cat_fwd_kernel(const T* __restrict__ dec, const T* __restrict__ s1, const T* __restrict__ s2,
                     int64_t sB, int64_t sC, T* __restrict__ cat, int Cd, int Cs, int64_t hw,
                     int64_t n_lerp, int64_t n_copy, int lerp_blocks) {
  constexpr int V = VECTOR ? Vec<T>::N : 1;
  const int Ct = Cd + Cs;
  if ((int)blockIdx.x < lerp_blocks) {
    const LerpW lw = lerp_weights();
    const int64_t hwv = hw / V;
    const bool small = n_lerp < (1ll << 31);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_lerp; i += (int64_t)lerp_blocks * 256) {
      int64_t bc, e, b, c;
      split_index(i, hwv, small, bc, e);
      e *= V;
      split_index(bc, Cs, small, b, c);
      const T* p1 = s1 + b * sB + c * sC + e;
      const T* p2 = s2 + b * sB + c * sC + e;
      T* o = cat + ((b * Ct + Cd + c) * 4) * hw + e;
      if constexpr (VECTOR) {
        const Vec<T> a = ldv(p1), bb = ldv(p2);
        Vec<T> m1, m2;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float fa = cvtf<T>(a.v[j]), fb = cvtf<T>(bb.v[j]);
          m1.v[j] = fromf<T>(fmaf(lw.a1, fa, __fmul_rn(lw.b1, fb)));
          m2.v[j] = fromf<T>(fmaf(lw.a2, fa, __fmul_rn(lw.b2, fb)));
        }
        stv(o, a); stv(o + hw, m1); stv(o + 2 * hw, m2); stv(o + 3 * hw, bb);
      } else {
        const float fa = ldf(p1), fb = ldf(p2);
        o[0] = *p1;
        o[hw] = fromf<T>(fmaf(lw.a1, fa, __fmul_rn(lw.b1, fb)));
        o[2 * hw] = fromf<T>(fmaf(lw.a2, fa, __fmul_rn(lw.b2, fb)));
        o[3 * hw] = *p2;
      }
    }
  } else {
    const int cb = (int)gridDim.x - lerp_blocks;
    const int64_t per_b = (int64_t)Cd * 4 * hw / V;   // vectors of dec per pair
    const bool small = n_copy < (1ll << 31);
    for (int64_t i = (int64_t)(blockIdx.x - lerp_blocks) * 256 + threadIdx.x; i < n_copy;
         i += (int64_t)cb * 256) {
      int64_t b, r;
      split_index(i, per_b, small, b, r);
      r *= V;
      const T* src = dec + b * (int64_t)Cd * 4 * hw + r;
      T* dst = cat + b * (int64_t)Ct * 4 * hw + r;
      if constexpr (VECTOR) stv(dst, ldv_stream(src));
      else *dst = *src;
    }
  }
}

template <typename T, bool VECTOR>
__global__ void __launch_bounds__(256)
tlerp_cat_bwd_kernel(const T* __restrict__ gcat, T* __restrict__ g1, T* __restrict__ g2, int64_t sB,
                     int64_t sC, int Cd, int Cs, int64_t hw, int64_t n) {
  constexpr int V = VECTOR ? Vec<T>::N : 1;
  const int Ct = Cd + Cs;
  const LerpW lw = lerp_weights();
  const int64_t hwv = hw / V;
  const bool small = n < (1ll << 31);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    int64_t bc, e, b, c;
    split_index(i, hwv, small, bc, e);
    e *= V;
    split_index(bc, Cs, small, b, c);
    const T* g = gcat + ((b * Ct + Cd + c) * 4) * hw + e;
    T* o1 = g1 + b * sB + c * sC + e;
    T* o2 = g2 + b * sB + c * sC + e;
    if constexpr (VECTOR) {
      const Vec<T> a = ldv_stream(g), m1 = ldv_stream(g + hw), m2 = ldv_stream(g + 2 * hw),
                   d = ldv_stream(g + 3 * hw);
      Vec<T> r1, r2;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float f0 = cvtf<T>(a.v[j]), f1 = cvtf<T>(m1.v[j]), f2 = cvtf<T>(m2.v[j]), f3 = cvtf<T>(d.v[j]);
        r1.v[j] = fromf<T>(fmaf(lw.a2, f2, fmaf(lw.a1, f1, f0)));
        r2.v[j] = fromf<T>(__fadd_rn(fmaf(lw.b2, f2, __fmul_rn(lw.b1, f1)), f3));
      }
      stv(o1, r1); stv(o2, r2);
    } else {
      const float f0 = cvtf<T>(g[0]), f1 = cvtf<T>(g[hw]), f2 = cvtf<T>(g[2 * hw]), f3 = cvtf<T>(g[3 * hw]);
      *o1 = fromf<T>(fmaf(lw.a2, f2, fmaf(lw.a1, f1, f0)));
      *o2 = fromf<T>(__fadd_rn(fmaf(lw.b2, f2, __fmul_rn(lw.b1, f1)), f3));
    }
  }
}


// ------------------------------------------------------------------------------
// channels-last (NDHWC) variants: memory order (B, T, h*w, C); vectors run along C
// ------------------------------------------------------------------------------
// One thread per 16-byte vector of the OUTPUT row (b, slot, pixel): consecutive threads write consecutive
// bytes of cat (the dec part, then the lerped skip part of the same pixel), so every 32-byte sector is completed
// by one warp even when Cd*sizeof(T) is not a multiple of 32 (SMOW_Net_LW's 28+16 channel level).  The skip
// vectors are re-read once per slot (4x, from L2/L1 — they are a small fraction of the concat).
// BatchNorm-apply (per-channel scale | shift, fp32 tensors only) + LeakyReLU of the decoder half
// per-channel parameter row `row` of a [rows][Cd] fp32 table for the N channels starting at c0 (c0 % 4 == 0): 16-byte loads
template <int N>
__device__ __forceinline__ void ld_params(float (&dst)[N], const float* __restrict__ table, int Cd, int row, int c0) {
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * Cd + c0 + j));
    dst[j] = t.x; dst[j + 1] = t.y; dst[j + 2] = t.z; dst[j + 3] = t.w;
  }
}
template <typename T>
__device__ __forceinline__ Vec<T> affine_leaky_vec(Vec<T> v, const float* __restrict__ affine, int Cd, int c0, float slope) {
  float sc[Vec<T>::N], sh[Vec<T>::N];
  ld_params<Vec<T>::N>(sc, affine, Cd, 0, c0);
  ld_params<Vec<T>::N>(sh, affine, Cd, 1, c0);
#pragma unroll
  for (int j = 0; j < Vec<T>::N; ++j) {
    const float f = fmaf(cvtf<T>(v.v[j]), sc[j], sh[j]);
    v.v[j] = fromf<T>(f > 0.f ? f : f * slope);
  }
  return v;
}

// LeakyReLU of the decoder half, fused into its copy (slope == 1: a plain bit-exact copy)
template <typename T>
__device__ __forceinline__ Vec<T> leaky_vec(Vec<T> v, float slope) {
  if (slope != 1.f) {
#pragma unroll
    for (int j = 0; j < Vec<T>::N; ++j) {
      const float f = cvtf<T>(v.v[j]);
      v.v[j] = fromf<T>(f > 0.f ? f : f * slope);
    }
  }
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
tlerp_cat_fwd_ndhwc_kernel(const T* __restrict__ dec, const T* __restrict__ s1, const T* __restrict__ s2,
                           int64_t sB, T* __restrict__ cat, int Cd, int Cs, int64_t hw, int64_t n_items,
                           bool do_copy, bool do_lerp, FastDiv fqt, FastDiv fhw, float slope,
                           const float* __restrict__ affine) {
  constexpr int V = Vec<T>::N;
  const int Ct = Cd + Cs;
  const int64_t qd = Cd / V, qt = Ct / V;
  const LerpW lw = lerp_weights();
  const bool small = n_items < (1ll << 31);
  if (small) {          // every index fits 32 bits: multiply-shift divisions, 32-bit offsets (the models' shapes)
    const uint32_t n32 = (uint32_t)n_items, step = gridDim.x * 256u, uqd = (uint32_t)qd;
    // The tensor is walked from its END: the producer of `dec` (frame mix / BatchNorm) wrote it front to back, so its tail is
    // what the L2 still holds; and the consumer of `cat` (a convolution starting at the front) finds our last writes there.
    for (uint32_t i0 = blockIdx.x * 256u + threadIdx.x; i0 < n32; i0 += step) {
      const uint32_t i = n32 - 1u - i0;
      uint32_t r, vt;
      fast_split(i, fqt, r, vt);               // r = (b*4 + slot)*hw + p
      T* dst = cat + ((int64_t)i * V);         // = cat + r*Ct + vt*V: the output is written densely
      if (vt < uqd) {
        if (do_copy) {
          const Vec<T> v = ldv_stream(dec + ((int64_t)r * Cd + vt * V));
          stv(dst, affine ? affine_leaky_vec(v, affine, Cd, (int)(vt * V), slope) : leaky_vec(v, slope));
        }
        continue;
      }
      if (!do_lerp) continue;
      uint32_t bs, p;
      fast_split(r, fhw, bs, p);
      const uint32_t slot = bs & 3u;
      const int64_t off = (int64_t)(bs >> 2) * sB + (int64_t)(p * (uint32_t)Cs + (vt - uqd) * V);
      Vec<T> o;
      if (slot == 0) o = ldv(s1 + off);
      else if (slot == 3) o = ldv(s2 + off);
      else {
        const Vec<T> a = ldv(s1 + off), bb = ldv(s2 + off);
        const float wa = slot == 1 ? lw.a1 : lw.a2, wb = slot == 1 ? lw.b1 : lw.b2;
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = fromf<T>(fmaf(wa, cvtf<T>(a.v[j]), __fmul_rn(wb, cvtf<T>(bb.v[j]))));
      }
      stv(dst, o);
    }
    return;
  }
  for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < n_items; i0 += (int64_t)gridDim.x * 256) {
    const int64_t i = n_items - 1 - i0;
    int64_t r, vt;
    split_index(i, qt, small, r, vt);          // r = (b*4 + slot)*hw + p
    if (vt < qd) {
      if (do_copy) {
        const Vec<T> v = ldv_stream(dec + r * Cd + vt * V);
        stv(cat + r * Ct + vt * V, affine ? affine_leaky_vec(v, affine, Cd, (int)(vt * V), slope) : leaky_vec(v, slope));
      }
      continue;
    }
    if (!do_lerp) continue;
    int64_t bs, p;
    split_index(r, hw, small, bs, p);
    const int64_t b = bs >> 2;
    const int slot = (int)(bs & 3);
    const int64_t off = b * sB + p * Cs + (vt - qd) * V;
    Vec<T> o;
    if (slot == 0) o = ldv(s1 + off);
    else if (slot == 3) o = ldv(s2 + off);
    else {
      const Vec<T> a = ldv(s1 + off), bb = ldv(s2 + off);
      const float wa = slot == 1 ? lw.a1 : lw.a2, wb = slot == 1 ? lw.b1 : lw.b2;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = fromf<T>(fmaf(wa, cvtf<T>(a.v[j]), __fmul_rn(wb, cvtf<T>(bb.v[j]))));
    }
    stv(cat + r * Ct + vt * V, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
tlerp_cat_bwd_ndhwc_kernel(const T* __restrict__ gcat, T* __restrict__ g1, T* __restrict__ g2, int64_t sB, int Cd,
                           int Cs, int64_t hw, int64_t n, FastDiv fqs, FastDiv fhw) {
  constexpr int V = Vec<T>::N;
  const int Ct = Cd + Cs;
  const LerpW lw = lerp_weights();
  const int64_t qs = Cs / V, fs = hw * Ct;
  const bool small = n < (1ll << 31) && (int64_t)4 * hw * Ct < (1ll << 31);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    int64_t bp, vs, b, p;
    if (small) {
      uint32_t ubp, uvs, ub, up;
      fast_split((uint32_t)i, fqs, ubp, uvs);
      fast_split(ubp, fhw, ub, up);
      bp = ubp; vs = uvs; b = ub; p = up;
    } else {
      split_index(i, qs, false, bp, vs);
      split_index(bp, hw, false, b, p);
    }
    const T* g = gcat + ((b * 4) * hw + p) * Ct + Cd + vs * V;
    const Vec<T> a = ldv_stream(g), m1 = ldv_stream(g + fs), m2 = ldv_stream(g + 2 * fs), d = ldv_stream(g + 3 * fs);
    Vec<T> r1, r2;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float f0 = cvtf<T>(a.v[j]), f1 = cvtf<T>(m1.v[j]), f2 = cvtf<T>(m2.v[j]), f3 = cvtf<T>(d.v[j]);
      r1.v[j] = fromf<T>(fmaf(lw.a2, f2, fmaf(lw.a1, f1, f0)));
      r2.v[j] = fromf<T>(__fadd_rn(fmaf(lw.b2, f2, __fmul_rn(lw.b1, f1)), f3));
    }
    stv(g1 + b * sB + p * Cs + vs * V, r1);
    stv(g2 + b * sB + p * Cs + vs * V, r2);
  }
}

// Backward of the fused activation + lerp + concat: blocks [0, nb_dec) turn the decoder half of the incoming gradient into
// d z = g * (z > 0 ? 1 : slope) (dense, one 16-byte vector per thread), the remaining blocks are the lerp backward above.
template <typename T>
__global__ void __launch_bounds__(256)
act_tlerp_cat_bwd_ndhwc_kernel(const T* __restrict__ gcat, const T* __restrict__ z, T* __restrict__ gz, T* __restrict__ g1,
                               T* __restrict__ g2, int64_t sB, int Cd, int Cs, int64_t hw, int64_t n_dec, int64_t n_skip,
                               int nb_dec, FastDiv fqd, FastDiv fqs, FastDiv fhw, float slope,
                               const float* __restrict__ bn) {
  // bn (fp32 tensors only, or null): [6][Cd] = scale | shift | mean | invstd | k1 = sum(du)/N | k2 = sum(du * xhat)/N.  With it
  // `z` is the block's PRE-BatchNorm tensor y and the dec branch is the whole BatchNorm + LeakyReLU backward:
  //   u = y*scale + shift, du = g * (u > 0 ? 1 : slope), xhat = (y - mean) * invstd, dy = scale * (du - k1 - xhat * k2)
  constexpr int V = Vec<T>::N;
  const int Ct = Cd + Cs;
  if ((int)blockIdx.x < nb_dec) {
    const int64_t qd = Cd / V;
    const bool small = n_dec < (1ll << 31);
    // walked from the END: the reduction pass over the same (gradient slice, y) ran front to back, its tail is L2-resident
    for (int64_t i0 = (int64_t)blockIdx.x * 256 + threadIdx.x; i0 < n_dec; i0 += (int64_t)nb_dec * 256) {
      const int64_t i = n_dec - 1 - i0;
      int64_t r, v;
      if (small) { uint32_t ur, uv; fast_split((uint32_t)i, fqd, ur, uv); r = ur; v = uv; }
      else split_index(i, qd, false, r, v);
      const Vec<T> g = ldv_stream(gcat + r * Ct + v * V), zz = ldv_stream(z + i * V);
      Vec<T> o;
      if (bn != nullptr) {
        const int c0 = (int)(v * V);
        float sc[V], sh[V], mu[V], is[V], k1[V], k2[V];
        ld_params<V>(sc, bn, Cd, 0, c0); ld_params<V>(sh, bn, Cd, 1, c0); ld_params<V>(mu, bn, Cd, 2, c0);
        ld_params<V>(is, bn, Cd, 3, c0); ld_params<V>(k1, bn, Cd, 4, c0); ld_params<V>(k2, bn, Cd, 5, c0);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float y = cvtf<T>(zz.v[j]);
          const float u = fmaf(y, sc[j], sh[j]);
          const float gg = cvtf<T>(g.v[j]), du = u > 0.f ? gg : gg * slope;
          const float xh = (y - mu[j]) * is[j];
          o.v[j] = fromf<T>(sc[j] * (du - k1[j] - xh * k2[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float gg = cvtf<T>(g.v[j]);
          o.v[j] = fromf<T>(cvtf<T>(zz.v[j]) > 0.f ? gg : gg * slope);
        }
      }
      stv(gz + i * V, o);
    }
    return;
  }
  if (n_skip == 0) return;
  const LerpW lw = lerp_weights();
  const int64_t qs = Cs / V, fs = hw * Ct;
  const int nb = gridDim.x - nb_dec;
  const bool small = n_skip < (1ll << 31) && (int64_t)4 * hw * Ct < (1ll << 31);
  for (int64_t i0 = (int64_t)(blockIdx.x - nb_dec) * 256 + threadIdx.x; i0 < n_skip; i0 += (int64_t)nb * 256) {
    const int64_t i = n_skip - 1 - i0;
    int64_t bp, vs, b, p;
    if (small) {
      uint32_t ubp, uvs, ub, up;
      fast_split((uint32_t)i, fqs, ubp, uvs);
      fast_split(ubp, fhw, ub, up);
      bp = ubp; vs = uvs; b = ub; p = up;
    } else {
      split_index(i, qs, false, bp, vs);
      split_index(bp, hw, false, b, p);
    }
    const T* g = gcat + ((b * 4) * hw + p) * Ct + Cd + vs * V;
    const Vec<T> a = ldv_stream(g), m1 = ldv_stream(g + fs), m2 = ldv_stream(g + 2 * fs), d = ldv_stream(g + 3 * fs);
    Vec<T> r1, r2;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float f0 = cvtf<T>(a.v[j]), f1 = cvtf<T>(m1.v[j]), f2 = cvtf<T>(m2.v[j]), f3 = cvtf<T>(d.v[j]);
      r1.v[j] = fromf<T>(fmaf(lw.a2, f2, fmaf(lw.a1, f1, f0)));
      r2.v[j] = fromf<T>(__fadd_rn(fmaf(lw.b2, f2, __fmul_rn(lw.b1, f1)), f3));
    }
    stv(g1 + b * sB + p * Cs + vs * V, r1);
    stv(g2 + b * sB + p * Cs + vs * V, r2);
  }
}

// Row-wise form of the BatchNorm + LeakyReLU + lerp backward (fp32, training — the form the models run).  ncu on the kernel
// above at LW's largest level (28 + 16 channels: 176-byte gradient rows): 478 MB read from HBM for 302 MB of operands, and
// the same whenever the decoder / skip boundary falls inside a 128-byte line (24 + 8 channels: 337 MB for 235 MB; 32 + 32:
// exact).  The decoder slice and the skip slice of a gradient row are read by different CTAs — or, in a first row-wise
// attempt, by different warps of one CTA whose loops drift apart — so by the time the second reader arrives the line has
// left the L2 (the gradient alone is larger than the cache) and comes from HBM again.
// Here ONE load instruction reads the whole row: thread = (pixel, 16-byte vector v of the concat row), 256 / q pixels per
// CTA iteration (q = (Cd + Cs) / 4); every thread loads its vector for the four frames (fully contiguous rows across the
// lanes), then lanes v < Cd / 4 do the BatchNorm + LeakyReLU backward of their vector (y loaded through L1, parameters in
// registers, 4 frames = 8 independent loads in flight) and the others the lerp backward (four frames in, two out).  Every
// gradient line is requested once.  Pixels are walked from the END (see above).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__global__ void __launch_bounds__(256, 3)
bn_act_tlerp_cat_bwd_rows_kernel(const float* __restrict__ gcat, const float* __restrict__ y, float* __restrict__ gy,
                                 float* __restrict__ g1, float* __restrict__ g2, int64_t sB, int Cd, int Cs, int64_t hw,
                                 int64_t npix, FastDiv fhw, float slope, const float* __restrict__ bn) {
  const int qd = Cd >> 2, q = (Cd + Cs) >> 2, ppi = 256 / q;
  const int pl = threadIdx.x / q, v = threadIdx.x - pl * q;
  if (pl >= ppi) return;
  const int Ct = Cd + Cs;
  const bool is_dec = v < qd;
  const int64_t fs = hw * Ct, fd = hw * Cd, step = (int64_t)gridDim.x * ppi;
  const bool small = npix < (1ll << 31);
  const int vp = is_dec ? v : 0;                                  // skip lanes load (and ignore) vector 0's parameters
  const float4 sc = ldg4(bn + 4 * vp), sh = ldg4(bn + Cd + 4 * vp), mu = ldg4(bn + 2 * Cd + 4 * vp);
  const float4 is = ldg4(bn + 3 * Cd + 4 * vp), k1 = ldg4(bn + 4 * Cd + 4 * vp), k2 = ldg4(bn + 5 * Cd + 4 * vp);
  const LerpW lw = lerp_weights();
  for (int64_t i0 = (int64_t)blockIdx.x * ppi + pl; i0 < npix; i0 += step) {
    const int64_t i = npix - 1 - i0;
    int64_t b, p;
    if (small) { uint32_t ub, up; fast_split((uint32_t)i, fhw, ub, up); b = ub; p = up; }
    else { b = i / hw; p = i - b * hw; }
    const int64_t r0 = b * 4 * hw + p;
    const float* g = gcat + r0 * Ct + v * 4;
    float4 gv[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) gv[f] = __ldcs(reinterpret_cast<const float4*>(g + f * fs));
    if (is_dec) {
      const float* yy = y + r0 * Cd + v * 4;
      float4 yv[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) yv[f] = ldg4(yy + f * fd);
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        float4 o;
#define SMOW_BN_BWD(c)                                                                   \
  {                                                                                      \
    const float u = fmaf(yv[f].c, sc.c, sh.c), du = u > 0.f ? gv[f].c : gv[f].c * slope; \
    const float xh = (yv[f].c - mu.c) * is.c;                                            \
    o.c = sc.c * (du - k1.c - xh * k2.c);                                                \
  }
        SMOW_BN_BWD(x) SMOW_BN_BWD(y) SMOW_BN_BWD(z) SMOW_BN_BWD(w)
#undef SMOW_BN_BWD
        *reinterpret_cast<float4*>(gy + (r0 + f * hw) * Cd + v * 4) = o;
      }
    } else {
      float4 r1, r2;
#define SMOW_LERP_BWD(c)                                                  \
  r1.c = fmaf(lw.a2, gv[2].c, fmaf(lw.a1, gv[1].c, gv[0].c));             \
  r2.c = __fadd_rn(fmaf(lw.b2, gv[2].c, __fmul_rn(lw.b1, gv[1].c)), gv[3].c);
      SMOW_LERP_BWD(x) SMOW_LERP_BWD(y) SMOW_LERP_BWD(z) SMOW_LERP_BWD(w)
#undef SMOW_LERP_BWD
      const int64_t o = b * sB + p * Cs + (v - qd) * 4;
      *reinterpret_cast<float4*>(g1 + o) = r1;
      *reinterpret_cast<float4*>(g2 + o) = r2;
    }
  }
}

// ------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------
template <typename T>
static int fwd_impl(const T* dec, const T* s1, const T* s2, int64_t sB, int64_t sC, T* cat, int B, int Cd,
                    int Cs, int64_t hw, cudaStream_t st) {
  constexpr int V = Vec<T>::N;
  const bool do_lerp = s1 != nullptr, do_copy = dec != nullptr && Cd > 0;
  if (!do_lerp && !do_copy) return 0;
  bool vec = (hw % V == 0) && aligned16(cat);
  if (do_lerp) vec = vec && aligned16(s1) && aligned16(s2) && (sB % V == 0) && (sC % V == 0);
  if (do_copy) vec = vec && aligned16(dec);
  const int v = vec ? V : 1;
  const int64_t n_lerp = do_lerp ? (int64_t)B * Cs * (hw / v) : 0;
  const int64_t n_copy = do_copy ? (int64_t)B * Cd * 4 * (hw / v) : 0;
  const int cap = device_info().sms * 8;
  int lb = (int)((n_lerp + 255) / 256 < cap ? (n_lerp + 255) / 256 : cap);
  int cb = (int)((n_copy + 511) / 512 < cap ? (n_copy + 511) / 512 : cap);
  if (vec)
    tlerp_cat_fwd_kernel<T, true><<<lb + cb, 256, 0, st>>>(dec, s1, s2, sB, sC, cat, Cd, Cs, hw, n_lerp, n_copy, lb);
  else
    tlerp_cat_fwd_kernel<T, false><<<lb + cb, 256, 0, st>>>(dec, s1, s2, sB, sC, cat, Cd, Cs, hw, n_lerp, n_copy, lb);
  count_launch();
  return check_launch("tlerp_cat_fwd");
}

template <typename T>
static int bwd_impl(const T* gcat, T* g1, T* g2, int64_t sB, int64_t sC, int B, int Cd, int Cs, int64_t hw,
                    cudaStream_t st) {
  constexpr int V = Vec<T>::N;
  const bool vec = (hw % V == 0) && aligned16(gcat) && aligned16(g1) && aligned16(g2) && (sB % V == 0) &&
                   (sC % V == 0);
  const int v = vec ? V : 1;
  const int64_t n = (int64_t)B * Cs * (hw / v);
  const int cap = device_info().sms * 8;
  const int nb = (int)((n + 255) / 256 < cap ? (n + 255) / 256 : cap);
  if (vec) tlerp_cat_bwd_kernel<T, true><<<nb, 256, 0, st>>>(gcat, g1, g2, sB, sC, Cd, Cs, hw, n);
  else tlerp_cat_bwd_kernel<T, false><<<nb, 256, 0, st>>>(gcat, g1, g2, sB, sC, Cd, Cs, hw, n);
  count_launch();
  return check_launch("tlerp_cat_bwd");
}

template <typename T>
static int fwd_ndhwc(const T* dec, const T* s1, const T* s2, int64_t sB, T* cat, int B, int Cd, int Cs, int64_t hw,
                     cudaStream_t st, float slope = 1.f, const float* affine = nullptr) {
  constexpr int V = Vec<T>::N;
  const bool do_lerp = s1 != nullptr, do_copy = dec != nullptr && Cd > 0;
  if (!do_lerp && !do_copy) return 0;
  if (Cd % V || Cs % V || !aligned16(cat) || (do_lerp && (!aligned16(s1) || !aligned16(s2) || sB % V)) ||
      (do_copy && !aligned16(dec)))
    return fail(SMOW_EALIGN, "tlerp_cat NDHWC needs Cd, Cs multiples of %d and 16 B aligned tensors", V);
  const int64_t n_items = (int64_t)B * 4 * hw * ((Cd + Cs) / V);
  const int cap = device_info().sms * 8;
  const int nb = (int)((n_items + 255) / 256 < cap ? (n_items + 255) / 256 : cap);
  tlerp_cat_fwd_ndhwc_kernel<T><<<nb, 256, 0, st>>>(dec, s1, s2, sB, cat, Cd, Cs, hw, n_items, do_copy, do_lerp,
                                                    make_fastdiv((Cd + Cs) / V), make_fastdiv(hw), slope, affine);
  count_launch();
  return check_launch("tlerp_cat_fwd (NDHWC)");
}

template <typename T>
static int act_bwd_ndhwc(const T* gcat, const T* z, T* gz, T* g1, T* g2, int64_t sB, int B, int Cd, int Cs, int64_t hw,
                         float slope, cudaStream_t st, const float* bn = nullptr) {
  constexpr int V = Vec<T>::N;
  if (Cd <= 0 || Cd % V || Cs % V || !aligned16(gcat) || !aligned16(z) || !aligned16(gz) ||
      (Cs > 0 && (!aligned16(g1) || !aligned16(g2) || sB % V)))
    return fail(SMOW_EALIGN, "act_tlerp_cat NDHWC needs Cd > 0, Cd and Cs multiples of %d and 16 B aligned tensors", V);
  if constexpr (std::is_same<T, float>::value) {
    const int q = (Cd + Cs) / 4;
    if (bn != nullptr && q <= 128 && option(OPT_BN_BWD_ROWS) != 0) {      // training form: whole concat rows per CTA
      const int ppi = 256 / q;
      const int64_t npix = (int64_t)B * hw, want = (npix + ppi - 1) / ppi;
      const int cap8 = device_info().sms * 3;                             // 3 resident CTAs per SM: one wave, static pixel split
      bn_act_tlerp_cat_bwd_rows_kernel<<<(int)(want < cap8 ? want : cap8), 256, 0, st>>>(gcat, z, gz, g1, g2, sB, Cd, Cs, hw, npix,
                                                                                         make_fastdiv(hw), slope, bn);
      count_launch();
      return check_launch("bn_act_tlerp_cat_bwd (rows)");
    }
  }
  const int64_t n_dec = (int64_t)B * 4 * hw * (Cd / V), n_skip = (int64_t)B * hw * (Cs / V);
  const int cap = device_info().sms * 8;
  const int nb_dec = (int)((n_dec + 255) / 256 < cap ? (n_dec + 255) / 256 : cap);
  const int nb_skip = (int)((n_skip + 255) / 256 < cap ? (n_skip + 255) / 256 : cap);
  act_tlerp_cat_bwd_ndhwc_kernel<T><<<nb_dec + nb_skip, 256, 0, st>>>(gcat, z, gz, g1, g2, sB, Cd, Cs, hw, n_dec, n_skip, nb_dec,
                                                                       make_fastdiv(Cd / V), make_fastdiv(Cs > 0 ? Cs / V : 1),
                                                                       make_fastdiv(hw), slope, bn);
  count_launch();
  return check_launch("act_tlerp_cat_bwd (NDHWC)");
}

template <typename T>
static int bwd_ndhwc(const T* gcat, T* g1, T* g2, int64_t sB, int B, int Cd, int Cs, int64_t hw, cudaStream_t st) {
  constexpr int V = Vec<T>::N;
  if (Cd % V || Cs % V || !aligned16(gcat) || !aligned16(g1) || !aligned16(g2) || sB % V)
    return fail(SMOW_EALIGN, "tlerp_cat NDHWC needs Cd, Cs multiples of %d and 16 B aligned tensors", V);
  const int64_t n = (int64_t)B * hw * (Cs / V);
  const int cap = device_info().sms * 8;
  const int nb = (int)((n + 255) / 256 < cap ? (n + 255) / 256 : cap);
  tlerp_cat_bwd_ndhwc_kernel<T><<<nb, 256, 0, st>>>(gcat, g1, g2, sB, Cd, Cs, hw, n, make_fastdiv(Cs / V), make_fastdiv(hw));
  count_launch();
  return check_launch("tlerp_cat_bwd (NDHWC)");
}

static int check_args(const void* cat, int B, int Cd, int Cs, int64_t hw, int dtype, int layout) {
  if (!cat) return fail(SMOW_EINVAL, "null pointer argument");
  if (B <= 0 || Cd < 0 || Cs <= 0 || hw <= 0) return fail(SMOW_EINVAL, "bad shape B=%d Cd=%d Cs=%d hw=%lld", B, Cd, Cs, (long long)hw);
  if (dtype != SMOW_F32 && dtype != SMOW_BF16) return fail(SMOW_EDTYPE, "unsupported dtype %d", dtype);
  if (layout != SMOW_NCDHW && layout != SMOW_NDHWC) return fail(SMOW_EDTYPE, "unsupported layout %d", layout);
  return 0;
}

}  // namespace smow

using namespace smow;

extern "C" {

int smow_tlerp_pair_cat_fwd(const void* dec, const void* skip_t1, const void* skip_t2, void* cat, int B, int Cd,
                            int Cs, int64_t hw, int dtype, int layout, void* stream) {
  if (int e = check_args(cat, B, Cd, Cs, hw, dtype, layout)) return e;
  if ((skip_t1 == nullptr) != (skip_t2 == nullptr)) return fail(SMOW_EINVAL, "skip_t1/skip_t2 must both be set or both be NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == SMOW_NDHWC) {
    if (dtype == SMOW_F32)
      return fwd_ndhwc<float>((const float*)dec, (const float*)skip_t1, (const float*)skip_t2, Cs * hw, (float*)cat, B,
                              Cd, Cs, hw, st);
    return fwd_ndhwc<__nv_bfloat16>((const __nv_bfloat16*)dec, (const __nv_bfloat16*)skip_t1,
                                    (const __nv_bfloat16*)skip_t2, Cs * hw, (__nv_bfloat16*)cat, B, Cd, Cs, hw, st);
  }
  if (dtype == SMOW_F32)
    return fwd_impl<float>((const float*)dec, (const float*)skip_t1, (const float*)skip_t2, Cs * hw, hw,
                           (float*)cat, B, Cd, Cs, hw, st);
  return fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)dec, (const __nv_bfloat16*)skip_t1,
                                 (const __nv_bfloat16*)skip_t2, Cs * hw, hw, (__nv_bfloat16*)cat, B, Cd, Cs, hw, st);
}

int smow_tlerp_cat_fwd(const void* dec, const void* skip, void* cat, int B, int Cd, int Cs, int64_t hw,
                       int dtype, int layout, void* stream) {
  if (int e = check_args(cat, B, Cd, Cs, hw, dtype, layout)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == SMOW_NDHWC) {   // stacked channels_last_3d: frame 2 starts hw*Cs elements after frame 1
    if (dtype == SMOW_F32) {
      const float* s = (const float*)skip;
      return fwd_ndhwc<float>((const float*)dec, s, s ? s + hw * Cs : nullptr, 2 * Cs * hw, (float*)cat, B, Cd, Cs, hw, st);
    }
    const __nv_bfloat16* s = (const __nv_bfloat16*)skip;
    return fwd_ndhwc<__nv_bfloat16>((const __nv_bfloat16*)dec, s, s ? s + hw * Cs : nullptr, 2 * Cs * hw,
                                    (__nv_bfloat16*)cat, B, Cd, Cs, hw, st);
  }
  if (dtype == SMOW_F32) {
    const float* s = (const float*)skip;
    return fwd_impl<float>((const float*)dec, s, s ? s + hw : nullptr, 2 * Cs * hw, 2 * hw, (float*)cat, B, Cd,
                           Cs, hw, st);
  }
  const __nv_bfloat16* s = (const __nv_bfloat16*)skip;
  return fwd_impl<__nv_bfloat16>((const __nv_bfloat16*)dec, s, s ? s + hw : nullptr, 2 * Cs * hw, 2 * hw,
                                 (__nv_bfloat16*)cat, B, Cd, Cs, hw, st);
}

int smow_act_tlerp_cat_fwd(const void* z, const void* skip_t1, const void* skip_t2, void* cat, int B, int Cd, int Cs,
                           int64_t hw, int64_t skip_pair_stride, float slope, int dtype, void* stream) {
  if (int e = check_args(cat, B, Cd, Cs, hw, dtype, SMOW_NDHWC)) return e;
  if (!z || !skip_t1 || !skip_t2 || Cd <= 0) return fail(SMOW_EINVAL, "act_tlerp_cat: null pointer / Cd == 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMOW_F32)
    return fwd_ndhwc<float>((const float*)z, (const float*)skip_t1, (const float*)skip_t2, skip_pair_stride, (float*)cat, B, Cd,
                            Cs, hw, st, slope);
  return fwd_ndhwc<__nv_bfloat16>((const __nv_bfloat16*)z, (const __nv_bfloat16*)skip_t1, (const __nv_bfloat16*)skip_t2,
                                  skip_pair_stride, (__nv_bfloat16*)cat, B, Cd, Cs, hw, st, slope);
}

int smow_act_tlerp_cat_bwd(const void* gcat, const void* z, void* gz, void* gskip_t1, void* gskip_t2, int B, int Cd, int Cs,
                           int64_t hw, int64_t skip_pair_stride, float slope, int dtype, void* stream) {
  if (int e = check_args(gcat, B, Cd, Cs, hw, dtype, SMOW_NDHWC)) return e;
  if (!z || !gz || !gskip_t1 || !gskip_t2) return fail(SMOW_EINVAL, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMOW_F32)
    return act_bwd_ndhwc<float>((const float*)gcat, (const float*)z, (float*)gz, (float*)gskip_t1, (float*)gskip_t2,
                                skip_pair_stride, B, Cd, Cs, hw, slope, st);
  return act_bwd_ndhwc<__nv_bfloat16>((const __nv_bfloat16*)gcat, (const __nv_bfloat16*)z, (__nv_bfloat16*)gz,
                                      (__nv_bfloat16*)gskip_t1, (__nv_bfloat16*)gskip_t2, skip_pair_stride, B, Cd, Cs, hw, slope, st);
}

// BatchNorm-apply + LeakyReLU + lerp + concat (fp32, NDHWC).  bn: [>= 2][Cd] = scale | shift (smow_bn_finalize).  Cs may be 0
// (no skip: the output is the activated block output itself, skip pointers NULL).
int smow_bn_act_tlerp_cat_fwd(const float* y, const float* bn, const float* skip_t1, const float* skip_t2, float* cat, int B,
                              int Cd, int Cs, int64_t hw, int64_t skip_pair_stride, float slope, void* stream) {
  if (!y || !bn || !cat || B <= 0 || Cd <= 0 || Cs < 0 || hw <= 0) return fail(SMOW_EINVAL, "bn_act_tlerp_cat: bad argument");
  if ((Cs > 0) != (skip_t1 != nullptr) || (skip_t1 == nullptr) != (skip_t2 == nullptr))
    return fail(SMOW_EINVAL, "bn_act_tlerp_cat: skip pointers must match Cs");
  if (!aligned16(bn)) return fail(SMOW_EALIGN, "bn_act_tlerp_cat: 16 B alignment");
  return fwd_ndhwc<float>(y, skip_t1, skip_t2, skip_pair_stride, cat, B, Cd, Cs, hw, (cudaStream_t)stream, slope, bn);
}

// second half of the backward (after smow_bn_act_bwd_reduce filled bn[4], bn[5]): gy dense + the lerp's gskip, one launch
int smow_bn_act_tlerp_cat_bwd(const float* gcat, const float* y, const float* bn, float* gy, float* gskip_t1, float* gskip_t2,
                              int B, int Cd, int Cs, int64_t hw, int64_t skip_pair_stride, float slope, void* stream) {
  if (!gcat || !y || !bn || !gy || B <= 0 || Cd <= 0 || Cs < 0 || hw <= 0) return fail(SMOW_EINVAL, "bn_act_tlerp_cat: bad argument");
  if (Cs > 0 && (!gskip_t1 || !gskip_t2)) return fail(SMOW_EINVAL, "bn_act_tlerp_cat: gskip pointers missing");
  if (!aligned16(bn)) return fail(SMOW_EALIGN, "bn_act_tlerp_cat: 16 B alignment");
  return act_bwd_ndhwc<float>(gcat, y, gy, gskip_t1, gskip_t2, skip_pair_stride, B, Cd, Cs, hw, slope, (cudaStream_t)stream, bn);
}

int smow_tlerp_pair_cat_bwd(const void* gcat, void* gskip_t1, void* gskip_t2, int B, int Cd, int Cs, int64_t hw,
                            int dtype, int layout, void* stream) {
  if (int e = check_args(gcat, B, Cd, Cs, hw, dtype, layout)) return e;
  if (!gskip_t1 || !gskip_t2) return fail(SMOW_EINVAL, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == SMOW_NDHWC) {
    if (dtype == SMOW_F32)
      return bwd_ndhwc<float>((const float*)gcat, (float*)gskip_t1, (float*)gskip_t2, Cs * hw, B, Cd, Cs, hw, st);
    return bwd_ndhwc<__nv_bfloat16>((const __nv_bfloat16*)gcat, (__nv_bfloat16*)gskip_t1, (__nv_bfloat16*)gskip_t2,
                                    Cs * hw, B, Cd, Cs, hw, st);
  }
  if (dtype == SMOW_F32)
    return bwd_impl<float>((const float*)gcat, (float*)gskip_t1, (float*)gskip_t2, Cs * hw, hw, B, Cd, Cs, hw, st);
  return bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)gcat, (__nv_bfloat16*)gskip_t1, (__nv_bfloat16*)gskip_t2,
                                 Cs * hw, hw, B, Cd, Cs, hw, st);
}

int smow_tlerp_cat_bwd(const void* gcat, void* gskip, int B, int Cd, int Cs, int64_t hw, int dtype, int layout,
                       void* stream) {
  if (int e = check_args(gcat, B, Cd, Cs, hw, dtype, layout)) return e;
  if (!gskip) return fail(SMOW_EINVAL, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (layout == SMOW_NDHWC) {
    if (dtype == SMOW_F32) {
      float* g = (float*)gskip;
      return bwd_ndhwc<float>((const float*)gcat, g, g + hw * Cs, 2 * Cs * hw, B, Cd, Cs, hw, st);
    }
    __nv_bfloat16* g = (__nv_bfloat16*)gskip;
    return bwd_ndhwc<__nv_bfloat16>((const __nv_bfloat16*)gcat, g, g + hw * Cs, 2 * Cs * hw, B, Cd, Cs, hw, st);
  }
  if (dtype == SMOW_F32) {
    float* g = (float*)gskip;
    return bwd_impl<float>((const float*)gcat, g, g + hw, 2 * Cs * hw, 2 * hw, B, Cd, Cs, hw, st);
  }
  __nv_bfloat16* g = (__nv_bfloat16*)gskip;
  return bwd_impl<__nv_bfloat16>((const __nv_bfloat16*)gcat, g, g + hw, 2 * Cs * hw, 2 * hw, B, Cd, Cs, hw, st);
}

}  // extern "C"
