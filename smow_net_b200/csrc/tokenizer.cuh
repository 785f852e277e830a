// Geometry shared by the two kernel families of the semantic tokenizer (row N2): tokenizer.cu (FP32-pipe kernels, any
// supported channel count) and tokenizer_mma.cu (tensor-core kernels, C = 16 / 32 — the two models' channel counts).
#pragma once
#include "common.cuh"

namespace smow {

constexpr int TOK_L = 8;            // token_len of both reference models

// chunk = pixels per CTA: its C*4-byte rows are one contiguous range of the stack, staged in shared memory by ONE bulk
// copy (cp.async.bulk + mbarrier) so that the stack leaves HBM exactly once per pass
struct TokGeom { int C, G, lpshift, chunk; int64_t hw; int nchunks; };      // LP = 4*G = 1 << lpshift lanes per pixel
__host__ __device__ inline int tok_chunk_px(int C) { return 16384 / C < 512 ? 16384 / C : 512; }   // <= 64 KB of x

// Fused form (rows A1 + N2): the tokenizer is the sole consumer of OFW's warped stack (models/SMOW_Net.py:50-52), so the
// chunk kernels can PRODUCE the staged rows themselves — frames 0 / 3 are bulk copies of the two input frames, frames 1 / 2
// are warped into shared memory with the reference's coordinate chain (common.cuh) — and the (B,C,4,H,W) stack never
// exists in HBM.  x: (B,2,HW,C) NDHWC; flow (B,2,2,H,W); xs / ys: the reference's linspace tables.
struct TokWarpSrc { const float* x; const float* flow; const float* xs; const float* ys; int H, W, wshift; };

// tensor-core family (tokenizer_mma.cu).  Both write the same per-chunk partials as the FP32-pipe kernels:
//   forward  part[bk][chunk] = [ m[8] | s[8] | T[8][C] ]      backward  part[bk][chunk] = [ dW[8][C] | db[8] ]
bool tok_mma_supported(int C);
int tok_mma_chunk_px(int C);
// `src` != nullptr selects the fused form (x is then unused).
void tok_fwd_mma_launch(const float* x, const TokWarpSrc* src, const float* wa, const float* ba, float* part,
                        const TokGeom& g, int B, cudaStream_t st);
void tok_bwd_mma_launch(const float* gtok, const float* x, const TokWarpSrc* src, const float* wa, const float* ba,
                        const float* tokens, const float* stats, float* gx, float* part, const TokGeom& g, int B,
                        cudaStream_t st);

}  // namespace smow
