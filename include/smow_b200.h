/*
 * smow_b200.h — C ABI of libsmow_b200.so (hand-written sm_100a CUDA kernels for
 * SMOW-Net's flow-guided bi-temporal alignment / fusion hot path).
 *
 * The reference has NO native seam for this path: it is Python calling ATen
 * (F.grid_sample / F.interpolate / torch.cat).  Each entry point below states
 * which reference lines it replaces (paths relative to the reference tree).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the
 *     current CUDA device of the calling thread; the caller owns every buffer;
 *   - nothing is allocated, nothing synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return 0 on success, a negative SMOW_E* code for an argument error and a
 *     positive value = cudaError_t for a launch failure.  A human-readable
 *     message for the last failure of the calling thread: smow_last_error();
 *   - no mutable global state except the launch counter, the tuning knobs and the monotonically increasing far-tap
 *     epoch of the warp backward (its stamp is stored in the CALLER's workspace: do not share one workspace block
 *     between launches that may be in flight on different streams), so calls are re-entrant from the forward
 *     thread and the autograd thread.
 *
 * Tensor vocabulary (the reference's own):
 *   frames   T1, T2   the two acquisition dates; a "pair" is one (T1,T2) sample
 *   x        bi-temporal feature stack  (B, C, 2, H, W)
 *   flow     predicted optical flow     (B, 2, 2, H, W)  = (b, {dx,dy}, frame, h, w), fp32
 *   out      temporal stack             (B, C, 4, H, W)  = [T1, warp(T1), warp(T2), T2]
 *   xs, ys   fp32 base-grid tables torch.linspace(-1,1,W) / (-1,1,H), built by the host
 *            exactly like models/SMOW_Net.py:617-618 so coordinates match bit-for-bit
 *
 * dtype : SMOW_F32 (0) or SMOW_BF16 (1) for feature tensors; flow, xs, ys and
 *         grad_flow are always fp32; all arithmetic is fp32.
 * layout: SMOW_NCDHW (0): features are contiguous (B,C,T,H,W) — what cuDNN hands
 *         the reference; SMOW_NDHWC (1): channels_last_3d, memory order
 *         (B,T,H,W,C).  flow is always contiguous (B,2,2,H,W).
 */
#ifndef SMOW_B200_H
#define SMOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMOW_ABI_VERSION 9

#if defined(__GNUC__)
#define SMOW_API __attribute__((visibility("default")))
#else
#define SMOW_API
#endif

enum { SMOW_F32 = 0, SMOW_BF16 = 1 };
enum { SMOW_NCDHW = 0, SMOW_NDHWC = 1 };

enum {
  SMOW_OK = 0,
  SMOW_EINVAL = -1,   /* bad shape / null pointer                 */
  SMOW_EALIGN = -2,   /* pointer or row pitch not 16-byte aligned  */
  SMOW_EDTYPE = -3,   /* unsupported dtype / layout combination    */
  SMOW_ERANGE = -4    /* tensor too large for 32-bit tile indexing */
};

SMOW_API int         smow_abi_version(void);
SMOW_API const char* smow_last_error(void);
/* Number of kernels this library has launched in this process (all threads). */
SMOW_API uint64_t    smow_launch_count(void);
/* Tuning knobs (ints), e.g. "warp_fwd_variant", "warp_bwd_variant".  Returns 0
 * or SMOW_EINVAL for an unknown key.  smow_get_option returns the value or -1. */
SMOW_API int         smow_set_option(const char* key, int value);
SMOW_API int         smow_get_option(const char* key);

/* ---- A1: OFW.flow_warp forward ------------------------------------------------
 * Replaces models/SMOW_Net.py:612-638 (= models/SMOW_Net_LW.py:454-480): base grid
 * + flow/[W,H], clamp(-1,1), per-frame grid_sample(bilinear, border,
 * align_corners=True) (ATen GridSampler.cuh:21-31,53-57), and the time-axis
 * concat [T1, warp(T1), warp(T2), T2] — one launch, warped frames never
 * round-trip HBM.
 *   x    (B,C,2,H,W)   out (B,C,4,H,W)                                            */
SMOW_API int smow_warp_stack_fwd(const void* x, const float* flow,
                        const float* xs, const float* ys, void* out,
                        int B, int C, int H, int W,
                        int dtype, int layout, void* stream);

/* Same, with the two frames given as separate (B,C,H,W) tensors (the Siamese
 * backbone outputs of models/SMOW_Net_LW.py:35-40, row A5): the unsqueeze+cat
 * that builds the stack is skipped.                                              */
SMOW_API int smow_warp_pair_fwd(const void* x_t1, const void* x_t2, const float* flow,
                       const float* xs, const float* ys, void* out,
                       int B, int C, int H, int W,
                       int dtype, int layout, void* stream);

/* ---- A1: backward of the same graph --------------------------------------------
 * Replaces the autograd chain CatBackward → GridSampler2DBackward (ATen
 * GridSampler.cuh:37-50,62-80) → ClampBackward → DivBackward → Select/Slice
 * backward:
 *   gx[:,:,0] = gout[:,:,0] + scatter(gout[:,:,1]);  gx[:,:,1] = gout[:,:,3] + scatter(gout[:,:,2])
 *   gflow     = d(out)/d(flow) with the clamp mask (inclusive at ±1) and the
 *               border-clip mask.  gx needs no zero-fill by the caller.
 *   gout (B,C,4,H,W)  x (B,C,2,H,W)  gx (B,C,2,H,W)  gflow (B,2,2,H,W) fp32
 * workspace: optional caller-owned device scratch (16-byte aligned; may be NULL).  The library still never
 *   allocates.  fp32 NDHWC only:
 *   - default backward (tile gather + far-tap pass): >= 64 bytes, zero-initialised ONCE by the caller and then
 *     reused by every call on the same stream (do not share one block between concurrent streams).  Word 0
 *     receives a per-call stamp when some source pixel has taps farther than one pixel away, so the far-tap
 *     pass can exit immediately for sub-pixel flows; without a workspace that pass re-reads the flow.
 *   - warp_bwd_variant = 3 (gather lists): smow_warp_bwd_workspace_bytes(B,H,W) bytes, uninitialised is fine.   */
SMOW_API int64_t smow_warp_bwd_workspace_bytes(int B, int H, int W);
SMOW_API int smow_warp_stack_bwd(const void* gout, const void* x, const float* flow,
                        const float* xs, const float* ys,
                        void* gx, float* gflow,
                        int B, int C, int H, int W,
                        int dtype, int layout, void* workspace, int64_t workspace_bytes, void* stream);

SMOW_API int smow_warp_pair_bwd(const void* gout, const void* x_t1, const void* x_t2,
                       const float* flow, const float* xs, const float* ys,
                       void* gx_t1, void* gx_t2, float* gflow,
                       int B, int C, int H, int W,
                       int dtype, int layout, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- A3+A4: temporal 2→4 lerp written straight into the decoder concat ----------
 * Replaces F.interpolate(xk, size=(4,h,w), 'trilinear', align_corners=True)
 * (models/SMOW_Net.py:64-73, models/SMOW_Net_LW.py:62-71) followed by
 * torch.cat([c3dtK, xK], dim=1) (models/SMOW_Net.py:78,82,86,90,94):
 *   cat[:, :Cd]        = dec                      (B,Cd,4,h,w)   (skipped if dec == NULL)
 *   cat[:, Cd:Cd+Cs]   = [T1, (1-l)T1+l*T2, (1-m)T1+m*T2, T2]    (skipped if skip == NULL)
 * with l = fp32(1/3), m = 2*l as ATen's upsample_trilinear3d computes them.
 * skip is (B,Cs,2,h,w) (or two (B,Cs,h,w) tensors for the _pair variant).
 * hw = h*w.  Cd may be 0 (stand-alone temporal upsample).                         */
SMOW_API int smow_tlerp_cat_fwd(const void* dec, const void* skip, void* cat,
                       int B, int Cd, int Cs, int64_t hw,
                       int dtype, int layout, void* stream);
SMOW_API int smow_tlerp_pair_cat_fwd(const void* dec, const void* skip_t1, const void* skip_t2,
                            void* cat, int B, int Cd, int Cs, int64_t hw,
                            int dtype, int layout, void* stream);

/* Backward: gskip[:,:,0] = g0 + (1-l)g1 + (1-m)g2 ; gskip[:,:,1] = l*g1 + m*g2 + g3
 * where g = gcat[:, Cd:Cd+Cs].  The dec half of gcat is consumed in place by the
 * caller as a strided view (what torch.cat's backward does), so it is not copied. */
SMOW_API int smow_tlerp_cat_bwd(const void* gcat, void* gskip,
                       int B, int Cd, int Cs, int64_t hw,
                       int dtype, int layout, void* stream);
SMOW_API int smow_tlerp_pair_cat_bwd(const void* gcat, void* gskip_t1, void* gskip_t2,
                            int B, int Cd, int Cs, int64_t hw,
                            int dtype, int layout, void* stream);

/* ---- A3+A4 with the decoder block's LeakyReLU folded in: NO copy of the decoder half ------------------------------------
 * The decoder half of every concat is the output of conv_trans_block_3d, whose last two operations are BatchNorm and
 * LeakyReLU(0.2) (models/SMOW_Net.py:136-137, models/SMOW_Net_LW.py:134-135).  Taking the BatchNorm output z instead,
 *   cat[:, :Cd]      = leaky_relu(z, slope)            (the activation pass the reference runs as its own kernel)
 *   cat[:, Cd:Cd+Cs] = [T1, (1-l)T1+l*T2, (1-m)T1+m*T2, T2]
 * one launch writes the whole concat buffer: the stand-alone LeakyReLU pass and the copy of smow_tlerp_cat_fwd are gone.
 * NDHWC only; skip frames as two pointers, `skip_pair_stride` elements between consecutive pairs (Cs*hw for separate
 * (B,Cs,h,w) tensors, 2*Cs*hw with skip_t2 = skip_t1 + Cs*hw for a stacked (B,Cs,2,h,w) tensor).
 * Backward: gz = gcat[:, :Cd] * (z > 0 ? 1 : slope) written dense, gskip as smow_tlerp_cat_bwd — one launch.        */
SMOW_API int smow_act_tlerp_cat_fwd(const void* z, const void* skip_t1, const void* skip_t2, void* cat,
                       int B, int Cd, int Cs, int64_t hw, int64_t skip_pair_stride, float slope, int dtype, void* stream);
SMOW_API int smow_act_tlerp_cat_bwd(const void* gcat, const void* z, void* gz, void* gskip_t1, void* gskip_t2,
                       int B, int Cd, int Cs, int64_t hw, int64_t skip_pair_stride, float slope, int dtype, void* stream);

/* ---- BatchNorm of the decoder blocks, folded into the kernels around it ---------------------------------------------------
 * conv_trans_block_3d.forward / conv_block_2_3d.forward end with  self.batch(mix) -> self.leaky  (models/SMOW_Net.py:135-137,
 * models/SMOW_Net_LW.py:133-135,173-175).  In training mode:
 *   1. smow_frame_mix_apply_tc_stats: the frame mix (above) + per-CTA partial sums (sum y, sum y^2 per channel) from its
 *      epilogue: stats [smow_frame_mix_stats_parts(B,C,T,hw)][2][C];
 *   2. smow_bn_finalize: bn[0] = scale = gamma*invstd, bn[1] = shift = beta - mean*scale, bn[2] = mean, bn[3] = invstd
 *      (bn is a [6][C] fp32 block); running_mean / running_var updated like nn.BatchNorm3d (momentum, unbiased variance);
 *   3. smow_bn_act_tlerp_cat_fwd: cat[:, :Cd] = leaky_relu(y*scale + shift), cat[:, Cd:] = lerped skip (Cs may be 0);
 *   backward: smow_bn_act_bwd_reduce (sum du, sum du*xhat -> bn[4], bn[5], dgamma, dbeta) then
 *   smow_bn_act_tlerp_cat_bwd (gy = scale*(du - k1 - xhat*k2) dense + the lerp's gskip).  fp32, NDHWC, deterministic.
 *   The apply pass reads every gradient row with one load instruction per (pixel, frame) — decoder slice and skip slice
 *   together — so each 128-byte line of the gradient comes from HBM once whatever the channel split (knob "bn_bwd_rows" = 0
 *   selects the earlier kernel with separate decoder / skip CTAs; identical results).                                    */
SMOW_API int     smow_frame_mix_stats_parts(int B, int C, int T, int64_t hw);
SMOW_API int     smow_frame_mix_apply_tc_stats(const float* in, const float* wpack, const float* bias, float* out, float* stats,
                         int B, int C, int T, int64_t hw, int64_t out_pitch, int shift, int own_off, void* stream);
SMOW_API int     smow_bn_finalize(const float* parts, int nparts, int C, int64_t count, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float momentum, float eps, float* bn, void* stream);
SMOW_API int     smow_bn_act_tlerp_cat_fwd(const float* y, const float* bn, const float* skip_t1, const float* skip_t2,
                         float* cat, int B, int Cd, int Cs, int64_t hw, int64_t skip_pair_stride, float slope, void* stream);
SMOW_API int64_t smow_bn_act_bwd_workspace_bytes(int B, int Cd, int64_t hw);
SMOW_API int     smow_bn_act_bwd_reduce(const float* gcat, const float* y, float* bn, float* dgamma, float* dbeta,
                         int B, int Cd, int Cs, int64_t hw, float slope, void* workspace, int64_t workspace_bytes, void* stream);
SMOW_API int     smow_bn_act_tlerp_cat_bwd(const float* gcat, const float* y, const float* bn, float* gy, float* gskip_t1,
                         float* gskip_t2, int B, int Cd, int Cs, int64_t hw, int64_t skip_pair_stride, float slope, void* stream);

/* ---- N2: semantic tokenizer (the sole consumer of the warped stack) ---------------------------
 * Replaces, per frame k of the stack, models/SMOW_Net.py:176-187 (= models/SMOW_Net_LW.py:195-206):
 *   spatial_attention = conv_a(x[:, :, k])  (1x1, C -> 8)  -> view(b, 8, H*W) -> softmax(dim=-1)
 *   tokens_k          = einsum('bln,bcn->blc', spatial_attention, x[:, :, k].view(b, c, H*W))
 * for all four frames in one pass over the stack (plus a small chunk-combine launch).  The
 * positional embedding and the concat of the four token sets stay with the caller.
 *   x       (B,C,4,H,W) fp32, layout SMOW_NDHWC only; C/4 a power of two <= 32
 *   wa      (8,C) = conv_a.weight[:, :, 0, 0]     ba (8) = conv_a.bias
 *   tokens  (B,4,8,C) out                          stats (B,4,16) out: per token max logit | 1/sum exp
 *   workspace: smow_tokenizer_workspace_bytes(B,C,hw) bytes, 16-byte aligned, uninitialised.
 * Backward: gx (B,C,4,H,W) NDHWC = d loss / d x through both the attention and the pooling,
 *   gwa (8,C), gba (8) overwritten (not accumulated); tokens / stats are the forward's outputs.
 * All sums run in a fixed order: results are bit-reproducible.
 * C = 16 / 32 (the two models): the three 8-token-wide GEMMs run on the tensor cores as warp-level TF32 MMAs
 * with a hi/lo split of both operands (3xTF32, fp32 accumulation): fp32-class accuracy (<= 4e-6 of scale
 * against a float64 evaluation, tests hold it to the same 1e-5 bar as the FP32-pipe kernels), ~2x their
 * speed.  Knob "tok_variant" = 0 selects the FP32-pipe kernels (the only family for other C).     */
SMOW_API int64_t smow_tokenizer_workspace_bytes(int B, int C, int64_t hw);
SMOW_API int smow_tokenizer_fwd(const void* x, const float* wa, const float* ba,
                       float* tokens, float* stats, int B, int C, int64_t hw,
                       int dtype, int layout, void* workspace, int64_t workspace_bytes, void* stream);
SMOW_API int smow_tokenizer_bwd(const float* gtokens, const void* x, const float* wa, const float* ba,
                       const float* tokens, const float* stats,
                       void* gx, float* gwa, float* gba, int B, int C, int64_t hw,
                       int dtype, int layout, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- A1 + N2 fused: the warped stack never reaches HBM -------------------------------------------
 * The tokenizer is the only consumer of OFW's output (models/SMOW_Net.py:50-52 -> :176-187; LW :48 -> :195-206), so
 * tokens = tokenizer(flow_warp(x, flow)) is ONE pass over the two input frames: every chunk kernel stages its pixel
 * rows itself — frames 0 / 3 are bulk copies of x[:, :, 0] / x[:, :, 1], frames 1 / 2 are warped into shared memory
 * with the arithmetic of smow_warp_stack_fwd (same coordinate chain, same tap order: the staged rows are bit-identical
 * to what that launch writes) — and feeds them to the tensor-core tokenizer.  HBM traffic forward: x read once per
 * frame it feeds (2 * 2*C*HW*4 B) + flow, instead of warp (6*C*HW*4) + tokenizer (4*C*HW*4).
 *   x (B,C,2,H,W) fp32 NDHWC, flow (B,2,2,H,W), xs / ys as for smow_warp_stack_fwd; C = 16 / 32 (the two models;
 *   smow_warp_tokenizer_supported(C) != 0, which also requires knob "tok_variant" != 0); tokens (B,4,8,C), stats (B,4,16),
 *   workspace: smow_tokenizer_workspace_bytes(B, C, H*W).
 * Backward: the staged rows are re-computed the same way (nothing was saved but x, flow, tokens, stats); gstack
 *   (B,C,4,H,W) NDHWC receives d loss / d stack — the `gout` of smow_warp_stack_bwd, which then yields gx and gflow;
 *   gwa (8,C), gba (8) overwritten.                                                                  */
SMOW_API int smow_warp_tokenizer_supported(int C);
SMOW_API int smow_warp_tokenizer_fwd(const void* x, const float* flow, const float* xs, const float* ys,
                       const float* wa, const float* ba, float* tokens, float* stats,
                       int B, int C, int H, int W, int dtype, int layout,
                       void* workspace, int64_t workspace_bytes, void* stream);
SMOW_API int smow_warp_tokenizer_bwd(const float* gtokens, const void* x, const float* flow,
                       const float* xs, const float* ys, const float* wa, const float* ba,
                       const float* tokens, const float* stats,
                       void* gstack, float* gwa, float* gba,
                       int B, int C, int H, int W, int dtype, int layout,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ---- N4: cyclic temporal frame mix of the decoder blocks -------------------------------------
 * Replaces, for C_in = C_out = C in {16, 28, 32, 64} (the large decoder levels), the slice / ten 1x1x1
 * convolutions / adds / concat of conv_trans_block_3d.forward and conv_block_2_3d.forward
 * (models/SMOW_Net.py:121-139, models/SMOW_Net_LW.py:121-137,160-175):
 *   out[b,f,p,:] = in[b,f,p,:] @ m0 + in[b,(f+shift)%4,p,:] @ m1[(f+own_off)%4]
 * in / out: (B,C,4,H,W) fp32 NDHWC; m0 (C,C) and m1 (4,C,C): row = input channel, column = output channel.
 *   forward  : m0 = W_shared, m1 = W_own,     shift = 1, own_off = 1
 *   d(input) : m0 = W_shared^T, m1 = W_own^T, shift = 3, own_off = 0   (in = d out)
 * smow_frame_mix_wgrad: gw (5,C,C) = dW_shared, dW_own[0..3] (overwritten); workspace of
 * smow_frame_mix_wgrad_workspace_bytes(B,C,hw) bytes, uninitialised.  Fixed summation order.             */
SMOW_API int     smow_frame_mix_supported(int C);
SMOW_API int     smow_frame_mix_apply(const float* in, const float* m0, const float* m1, float* out,
                         int B, int C, int64_t hw, int shift, int own_off, void* stream);
SMOW_API int64_t smow_frame_mix_wgrad_workspace_bytes(int B, int C, int64_t hw);
SMOW_API int     smow_frame_mix_wgrad(const float* x, const float* gy, float* gw, int B, int C, int64_t hw,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ---- N4 on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in tensor memory, TMA tile loads
 * and stores).  Same operation as smow_frame_mix_apply, generalised:
 *   out[b,f,p,0:C] = in[b,f,p,:] @ W_0 + in[b,(f+shift)%T,p,:] @ W_{1+(f+own_off)%T}  (+ bias[f,:])
 *   T = 4: decoder blocks (models/SMOW_Net.py:121-139, models/SMOW_Net_LW.py:119-137,160-175);
 *   T = 2: the encoder's Decompose_conv temporal exchange (models/SMOW_Net.py:460-473).
 *   in     (B,C,T,H,W) fp32 NDHWC, dense
 *   wpack  (1+T, C_out, C_in) fp32: matrix m as [output channel][input channel] (a Conv3d 1x1x1 weight as stored;
 *          a ConvTranspose3d weight transposed); for d(input) pass the transposed matrices, shift = T-1, own_off = 0
 *   bias   (T, C) or NULL
 *   out    rows of `out_pitch` floats per pixel (out_pitch >= C, multiple of 4): out_pitch > C writes the result
 *          straight into channels [0,C) of the decoder's concat buffer (rows A3+A4: models/SMOW_Net.py:78-94), the
 *          buffer smow_tlerp_cat_fwd then completes with dec == NULL.
 *   C in {16, 28, 32k (k <= 8), 320, 384, 448, 512}.  Arithmetic: TF32 products (10-bit mantissa), fp32 accumulation —
 *   what cuDNN runs the reference's 1x1x1 convolutions in under torch.backends.cudnn.allow_tf32 = True.        */
SMOW_API int smow_frame_mix_tc_supported(int C, int T);
SMOW_API int smow_frame_mix_apply_tc(const float* in, const float* wpack, const float* bias, float* out,
                         int B, int C, int T, int64_t hw, int64_t out_pitch, int shift, int own_off, void* stream);
/* Weight gradients of the same operation on the tensor cores (M = input channels, N = output channels, K = pixels;
 * the NDHWC tiles are MN-major operands as TMA delivers them, nothing is transposed):
 *   gw[0]     = sum_{b,f} x[b,f]^T gy[b,f]                      (dW_0, row = input channel, column = output channel)
 *   gw[1+g]   = sum_b x[b,(f+shift)%T]^T gy[b,f],  g = (f+own_off)%T
 * x, gy (B,C,T,H,W) fp32 NDHWC dense; gw (1+T, C, C) overwritten; workspace of
 * smow_frame_mix_wgrad_tc_workspace_bytes(B,C,T,hw) bytes (per-CTA partials, added in index order: deterministic). */
SMOW_API int64_t smow_frame_mix_wgrad_tc_workspace_bytes(int B, int C, int T, int64_t hw);
SMOW_API int smow_frame_mix_wgrad_tc(const float* x, const float* gy, float* gw, int B, int C, int T, int64_t hw,
                         int shift, int own_off, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- N1: the OFW flow head -----------------------------------------------------------------------------------------
 * Replaces models/SMOW_Net.py:606-608 (= models/SMOW_Net_LW.py:448-450):
 *     seg_up = F.interpolate(seg_down, size=(2,H,W), 'trilinear', align_corners=True)
 *     flow   = flow_make(torch.cat([x, seg_up], 1))            Conv3d(2C -> 2, 3x3x3, padding 1, no bias)
 * with neither the up-sampled tensor nor the concat:  flow = stencil_x(x; W[:, :C]) + sum_tap bilerp(Z[tap], p + tap), where
 *     Z[b, tap, i, j, 2t+o] = sum_{t', c} W[o, C+c, t'-t+1, kh, kw] * seg_down[b, c, t', i, j]      (tap = 3 kh + kw)
 * is the channel contraction done at LOW resolution by the caller (a tiny einsum; up-sampling and contraction commute).
 *   x (B,C,2,H,W) fp32 NDHWC; weight (2,2C,3,3,3) = flow_make.weight; z (B,9,h,w,4); flow (B,2,2,H,W) contiguous out.
 * Backward: gx (B,C,2,H,W) NDHWC, gweight (2,2C,3,3,3) — ONLY the x half [:, :C] is written, gz (B,9,h,w,4); the caller
 *   back-propagates gz through its einsum to seg_down and W[:, C:].  Deterministic (no atomics).
 * workspace: smow_flow_head_workspace_bytes(B,C,H,W) bytes, uninitialised.  C in {16, 32, 64}, W % 4 == 0.          */
SMOW_API int     smow_flow_head_supported(int C, int H, int W, int h, int w);
SMOW_API int64_t smow_flow_head_workspace_bytes(int B, int C, int H, int W);
SMOW_API int     smow_flow_head_fwd(const float* x, const float* weight, const float* z, float* flow,
                         int B, int C, int H, int W, int h, int w, void* workspace, int64_t workspace_bytes, void* stream);
SMOW_API int     smow_flow_head_bwd(const float* gflow, const float* x, const float* weight, float* gx, float* gweight,
                         float* gz, int B, int C, int H, int W, int h, int w,
                         void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMOW_B200_H */
