/*
 * smow_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, fp32 CPU restatement of the reference algorithm for SMOW-Net's flow-guided
 * alignment/fusion path.  It exists to check the CUDA kernels (tests/, smoke()) and to
 * serve as the timed CPU baseline of bench.py; the product (smow_net_b200/) never links,
 * imports or calls it.  Parity status: PINNED — checked against outputs of the reference's
 * own code (OFW.flow_warp + autograd, F.interpolate + torch.cat) generated in the build
 * container by oracle/make_golden.py and committed under tests/golden/.
 *
 * Each function cites the reference lines it follows (paths relative to the reference
 * tree; "GridSampler.cuh" is ATen's aten/src/ATen/native/cuda/GridSampler.cuh, the
 * third-party file where the arithmetic of F.grid_sample lives).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).
 * -ffp-contract=off keeps every rounding step explicit, like the reference's op-by-op
 * PyTorch graph.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define API __attribute__((visibility("default")))

/* models/SMOW_Net.py:627-631: g = base + flow/size, clamp(-1,1);
 * GridSampler.cuh:21-31 (unnormalize, align_corners): ((g+1)/2)*(size-1);
 * GridSampler.cuh:53-57 (clip_coordinates) and :62-80 (its gradient gate). */
typedef struct {
  float i;      /* clipped source coordinate, pixels */
  float gate;   /* 1 if d(i)/d(flow) is non-zero (clamp mask inclusive, clip mask exclusive) */
} axis_t;

static axis_t axis_coord(float base, float flow, int size) {
  axis_t a;
  float g = base + flow / (float)size;
  int pass_clamp = (g >= -1.0f) && (g <= 1.0f);
  if (g < -1.0f) g = -1.0f;
  if (g > 1.0f) g = 1.0f;
  float i = ((g + 1.0f) / 2.0f) * (float)(size - 1);
  int pass_clip = (i > 0.0f) && (i < (float)(size - 1));
  if (i < 0.0f) i = 0.0f;
  if (i > (float)(size - 1)) i = (float)(size - 1);
  a.i = i;
  a.gate = (pass_clamp && pass_clip) ? 1.0f : 0.0f;
  return a;
}

/* element (b,c,t,h,w) of a contiguous (B,C,T,H,W) tensor */
#define IDX5(b, c, t, p, C, T, HW) ((((size_t)(b) * (C) + (c)) * (T) + (t)) * (HW) + (p))

/* OFW.flow_warp forward — models/SMOW_Net.py:612-638.
 * x (B,C,2,H,W), flow (B,2,2,H,W), xs = linspace(-1,1,W), ys = linspace(-1,1,H),
 * out (B,C,4,H,W) = [x_T1, warp(x_T1,flow_T1), warp(x_T2,flow_T2), x_T2]. */
API void smow_oracle_warp_stack_fwd(const float* x, const float* flow, const float* xs, const float* ys,
                                    float* out, int B, int C, int H, int W) {
  const int HW = H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < 2; ++t)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const int p = h * W + w;
          const float fx = flow[IDX5(b, 0, t, p, 2, 2, HW)];
          const float fy = flow[IDX5(b, 1, t, p, 2, 2, HW)];
          const axis_t ax = axis_coord(xs[w], fx, W), ay = axis_coord(ys[h], fy, H);
          /* GridSampler.cu bilinear: corners and weights nw, ne, sw, se */
          const float x0f = floorf(ax.i), y0f = floorf(ay.i);
          const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
          const float nw = (x0f + 1.0f - ax.i) * (y0f + 1.0f - ay.i);
          const float ne = (ax.i - x0f) * (y0f + 1.0f - ay.i);
          const float sw = (x0f + 1.0f - ax.i) * (ay.i - y0f);
          const float se = (ax.i - x0f) * (ay.i - y0f);
          for (int c = 0; c < C; ++c) {
            const float* pl = x + IDX5(b, c, t, 0, C, 2, HW);
            float acc = pl[y0 * W + x0] * nw;               /* (y0,x0) is always in bounds after the clip */
            if (x1 <= W - 1) acc += pl[y0 * W + x1] * ne;   /* out-of-bounds taps are skipped (:219-222) */
            if (y1 <= H - 1) acc += pl[y1 * W + x0] * sw;
            if (x1 <= W - 1 && y1 <= H - 1) acc += pl[y1 * W + x1] * se;
            out[IDX5(b, c, 1 + t, p, C, 4, HW)] = acc;
            out[IDX5(b, c, t ? 3 : 0, p, C, 4, HW)] = pl[p];   /* models/SMOW_Net.py:634-636 */
          }
        }
}

/* Backward of the same graph:
 *   gx[:,:,t]  = gout[:,:,pass(t)] + sum over output pixels of weight * gout[:,:,1+t]   (safe_add_2d, :248-262)
 *   gflow      = gix_mult * gix, masked by the clamp (inclusive) and divided by [W,H]. */
API void smow_oracle_warp_stack_bwd(const float* gout, const float* x, const float* flow, const float* xs,
                                    const float* ys, float* gx, float* gflow, int B, int C, int H, int W) {
  const int HW = H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < 2; ++t) {
      for (int c = 0; c < C; ++c) {   /* pass-through slots (SliceBackward) */
        const float* g = gout + IDX5(b, c, t ? 3 : 0, 0, C, 4, HW);
        float* o = gx + IDX5(b, c, t, 0, C, 2, HW);
        for (int p = 0; p < HW; ++p) o[p] = g[p];
      }
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          const int p = h * W + w;
          const float fx = flow[IDX5(b, 0, t, p, 2, 2, HW)];
          const float fy = flow[IDX5(b, 1, t, p, 2, 2, HW)];
          const axis_t ax = axis_coord(xs[w], fx, W), ay = axis_coord(ys[h], fy, H);
          const float x0f = floorf(ax.i), y0f = floorf(ay.i);
          const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
          const float wx0 = x0f + 1.0f - ax.i, wx1 = ax.i - x0f;
          const float wy0 = y0f + 1.0f - ay.i, wy1 = ay.i - y0f;
          const int x1ok = x1 <= W - 1, y1ok = y1 <= H - 1;
          float gix = 0.0f, giy = 0.0f;
          for (int c = 0; c < C; ++c) {
            const float go = gout[IDX5(b, c, 1 + t, p, C, 4, HW)];
            const float* pl = x + IDX5(b, c, t, 0, C, 2, HW);
            float* gp = gx + IDX5(b, c, t, 0, C, 2, HW);
            {
              const float v = pl[y0 * W + x0];
              gp[y0 * W + x0] += (wx0 * wy0) * go;
              gix -= v * wy0 * go; giy -= v * wx0 * go;
            }
            if (x1ok) {
              const float v = pl[y0 * W + x1];
              gp[y0 * W + x1] += (wx1 * wy0) * go;
              gix += v * wy0 * go; giy -= v * wx1 * go;
            }
            if (y1ok) {
              const float v = pl[y1 * W + x0];
              gp[y1 * W + x0] += (wx0 * wy1) * go;
              gix -= v * wy1 * go; giy += v * wx0 * go;
            }
            if (x1ok && y1ok) {
              const float v = pl[y1 * W + x1];
              gp[y1 * W + x1] += (wx1 * wy1) * go;
              gix += v * wy1 * go; giy += v * wx1 * go;
            }
          }
          /* GridSampler.cuh:37-50: d(unnormalize) = (size-1)/2 ; ClampBackward ; DivBackward by [W,H] */
          gflow[IDX5(b, 0, t, p, 2, 2, HW)] = (ax.gate * ((float)(W - 1) / 2.0f) * gix) / (float)W;
          gflow[IDX5(b, 1, t, p, 2, 2, HW)] = (ay.gate * ((float)(H - 1) / 2.0f) * giy) / (float)H;
        }
    }
}

/* Temporal 2->4 upsample + decoder concat — models/SMOW_Net.py:64-73 and :78,82,86,90,94.
 * ATen upsample_trilinear3d(align_corners=True): rdepth = (2-1)/(4-1) in fp32,
 * t1r = rdepth*t2, t1 = (int)t1r, lambda1 = t1r - t1, lambda0 = 1 - lambda1; spatial size is
 * unchanged so the h/w lambdas are (1,0).  dec may be NULL (Cd = 0). */
API void smow_oracle_tlerp_cat_fwd(const float* dec, const float* skip, float* cat, int B, int Cd, int Cs,
                                   int64_t hw) {
  const float rdepth = (float)(2 - 1) / (float)(4 - 1);
  const int Ct = Cd + Cs;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int64_t i = 0; i < (int64_t)Cd * 4 * hw; ++i) cat[(size_t)b * Ct * 4 * hw + i] = dec[(size_t)b * Cd * 4 * hw + i];
    for (int c = 0; c < Cs; ++c)
      for (int t2 = 0; t2 < 4; ++t2) {
        const float t1r = rdepth * (float)t2;
        const int t1 = (int)t1r;
        const int t1p = (t1 < 1) ? 1 : 0;
        const float l1 = t1r - (float)t1, l0 = 1.0f - l1;
        const float* a = skip + IDX5(b, c, t1, 0, Cs, 2, hw);
        const float* bb = skip + IDX5(b, c, t1 + t1p, 0, Cs, 2, hw);
        float* o = cat + IDX5(b, Cd + c, t2, 0, Ct, 4, hw);
        for (int64_t p = 0; p < hw; ++p) o[p] = l0 * a[p] + l1 * bb[p];
      }
  }
}

API void smow_oracle_tlerp_cat_bwd(const float* gcat, float* gskip, int B, int Cd, int Cs, int64_t hw) {
  const float rdepth = (float)(2 - 1) / (float)(4 - 1);
  const int Ct = Cd + Cs;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < Cs; ++c) {
      float* g0 = gskip + IDX5(b, c, 0, 0, Cs, 2, hw);
      float* g1 = gskip + IDX5(b, c, 1, 0, Cs, 2, hw);
      for (int64_t p = 0; p < hw; ++p) { g0[p] = 0.0f; g1[p] = 0.0f; }
      for (int t2 = 0; t2 < 4; ++t2) {
        const float t1r = rdepth * (float)t2;
        const int t1 = (int)t1r;
        const int t1p = (t1 < 1) ? 1 : 0;
        const float l1 = t1r - (float)t1, l0 = 1.0f - l1;
        const float* g = gcat + IDX5(b, Cd + c, t2, 0, Ct, 4, hw);
        float* lo = gskip + IDX5(b, c, t1, 0, Cs, 2, hw);
        float* hi = gskip + IDX5(b, c, t1 + t1p, 0, Cs, 2, hw);
        for (int64_t p = 0; p < hw; ++p) { lo[p] += l0 * g[p]; hi[p] += l1 * g[p]; }
      }
    }
}

/* ---- row N2: semantic tokenizer --------------------------------------------------------------
 * models/SMOW_Net.py:176-187 (= models/SMOW_Net_LW.py:195-206), per pair b and frame k:
 *   spatial_attention = conv_a(x_t)            1x1 conv C -> L:  logit[l][p] = sum_c wa[l][c] x[c][p] + ba[l]
 *   softmax over the H*W pixels (dim = -1)      ATen: max, exp(x - max), sum, divide
 *   tokens = einsum('bln,bcn->blc')            tokens[l][c] = sum_p attn[l][p] x[c][p]
 * x (B,C,4,hw) NCDHW contiguous; tokens (B,4,L,C).  Double accumulation: this is the checker. */
API void smow_oracle_tokenizer_fwd(const float* x, const float* wa, const float* ba, float* tokens, int B, int C,
                                   int L, int64_t hw) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < L; ++l) {
        const float* xb = x + (((int64_t)b * C) * 4 + k) * hw;      /* channel c at + c*4*hw */
        double mx = -INFINITY;
        for (int64_t p = 0; p < hw; ++p) {
          double lg = ba[l];
          for (int c = 0; c < C; ++c) lg += (double)wa[l * C + c] * xb[(int64_t)c * 4 * hw + p];
          if (lg > mx) mx = lg;
        }
        double s = 0.0;
        double acc[512];                                            /* C <= 512, checked by the Python front-end */
        for (int c = 0; c < C; ++c) acc[c] = 0.0;
        for (int64_t p = 0; p < hw; ++p) {
          double lg = ba[l];
          for (int c = 0; c < C; ++c) lg += (double)wa[l * C + c] * xb[(int64_t)c * 4 * hw + p];
          const double e = exp(lg - mx);
          s += e;
          for (int c = 0; c < C; ++c) acc[c] += e * xb[(int64_t)c * 4 * hw + p];
        }
        for (int c = 0; c < C; ++c) tokens[(((int64_t)b * 4 + k) * L + l) * C + c] = (float)(acc[c] / s);
      }
}

/* ---- row N4: cyclic temporal frame mix -------------------------------------------------------
 * models/SMOW_Net.py:121-139 (= models/SMOW_Net_LW.py:119-137, 160-175):
 *   x = cat([T1_F1 + T2_F2, T2_F1 + T3_F2, T3_F1 + T4_F2, T4_F1 + T1_F2], dim=2)
 * with Tk_F1 = conv3d_time_5(Tk) (shared) and Tk_F2 = conv3d_time_k(Tk) (own), all 1x1x1:
 *   out[b][d][j][p] = sum_c x[b][c][j][p] ws[c][d] + sum_c x[b][c][(j+1)%4][p] wo[(j+1)%4][c][d] (+ bias[j][d])
 * x (B,Cin,4,hw), out (B,Cout,4,hw) NCDHW contiguous; ws (Cin,Cout), wo (4,Cin,Cout): rows = input channels;
 * bias (4,Cout) or NULL. */
API void smow_oracle_frame_mix_fwd(const float* x, const float* ws, const float* wo, const float* bias, float* out,
                                   int B, int Cin, int Cout, int64_t hw) {
#pragma omp parallel for collapse(2)
  for (int b = 0; b < B; ++b)
    for (int d = 0; d < Cout; ++d)
      for (int j = 0; j < 4; ++j) {
        const int k = (j + 1) % 4;
        for (int64_t p = 0; p < hw; ++p) {
          double a = bias ? (double)bias[j * Cout + d] : 0.0;
          for (int c = 0; c < Cin; ++c) {
            a += (double)x[(((int64_t)b * Cin + c) * 4 + j) * hw + p] * ws[c * Cout + d];
            a += (double)x[(((int64_t)b * Cin + c) * 4 + k) * hw + p] * wo[((int64_t)k * Cin + c) * Cout + d];
          }
          out[(((int64_t)b * Cout + d) * 4 + j) * hw + p] = (float)a;
        }
      }
}
