// A1, tiled variants (NCDHW): persistent CTAs walk (pair, frame, row-band) tiles and
// pipeline channel chunks of each tile through shared memory with 1-D bulk copies
// (cp.async.bulk + mbarrier).  A tile spans whole rows, so one channel of a tile is
// one contiguous byte range in HBM and in the (B,C,4,H,W) output.
//
// Forward: per tile the 4 bilinear weights / tap offsets of every pixel are computed
// once and kept in registers; per channel chunk the four taps come from the staged
// tile (+/-HALO rows), the un-warped T1/T2 slot leaves shared memory as a bulk store
// and the warped slot as coalesced 128 B stores.  Taps farther than HALO rows away
// (large flows) are fetched from global memory, so any displacement stays correct.
#pragma once
#include "common.cuh"
#include "bulk.cuh"

namespace smow {

// ------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------
template <typename T, int NP, int CC, int NSTAGE>
__global__ void __launch_bounds__(256)
warp_fwd_tiled_kernel(const T* __restrict__ x1, const T* __restrict__ x2, int64_t sB, int64_t sC,
                      const float* __restrict__ flow, const float* __restrict__ xs,
                      const float* __restrict__ ys, T* __restrict__ out,
                      int C, int H, int W, int R, int HALO, int nbands, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int WR = R + 2 * HALO;
  const int plane = WR * W;                       // elements of one channel in a stage
  const uint32_t stage_bytes = (uint32_t)(CC * plane * sizeof(T));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NSTAGE * stage_bytes);
  const int tid = threadIdx.x;
  const int HW = H * W;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(full + s, 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int nchunk = C / CC;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_it = my_tiles * nchunk;

  auto issue = [&](int it) {  // thread 0: enqueue the bulk loads of pipeline iteration `it`
    const int tl = it / nchunk, ch = it - tl * nchunk;
    const int tile = blockIdx.x + tl * gridDim.x;
    const int band = tile % nbands, bt = tile / nbands;
    const int b = bt >> 1, t = bt & 1;
    const int h0 = band * R;
    const int r_lo = max(0, h0 - HALO), r_hi = min(H, h0 + R + HALO);
    const uint32_t bytes = (uint32_t)((r_hi - r_lo) * W * sizeof(T));
    const int s = it % NSTAGE;
    T* dst = reinterpret_cast<T*>(smem_raw + (size_t)s * stage_bytes) + (r_lo - (h0 - HALO)) * W;
    const T* src = (t ? x2 : x1) + b * sB + (int64_t)(ch * CC) * sC + r_lo * W;
    mbar_expect_tx(full + s, bytes * CC);
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) bulk_g2s(dst + cc * plane, src + cc * sC, bytes, full + s);
  };
  if (tid == 0) {
    const int pre = total_it < NSTAGE ? total_it : NSTAGE;
    for (int it = 0; it < pre; ++it) issue(it);
  }

  // per-pixel state of the current tile (registers)
  float w_nw[NP], w_ne[NP], w_sw[NP], w_se[NP];
  int soff[NP];   // smem offset of the nw tap inside a channel plane, or -1: fetch from global
  int goff[NP];   // global offset of the nw tap inside a channel plane
  int ooff[NP];   // output pixel offset h*W+w, or -1 for a masked lane
  int tb = 0, tt = 0, th0 = 0;

  for (int it = 0; it < total_it; ++it) {
    const int tl = it / nchunk, ch = it - tl * nchunk;
    if (ch == 0) {
      const int tile = blockIdx.x + tl * gridDim.x;
      const int band = tile % nbands, bt = tile / nbands;
      tb = bt >> 1; tt = bt & 1; th0 = band * R;
      const float* fl = flow + ((int64_t)(tb * 2) * 2 + tt) * HW;
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int pl = tid + k * 256;
        const int r = pl / W, col = pl - r * W;
        const int h = th0 + r;
        if (r < R && h < H) {
          const int p = h * W + col;
          const Footprint fp = footprint(__ldg(xs + col), __ldg(ys + h), __ldg(fl + p),
                                         __ldg(fl + p + 2 * (int64_t)HW), W, H);
          w_nw[k] = __fmul_rn(fp.wx0, fp.wy0);
          w_ne[k] = fp.x1ok ? __fmul_rn(fp.wx1, fp.wy0) : 0.f;
          w_sw[k] = fp.y1ok ? __fmul_rn(fp.wx0, fp.wy1) : 0.f;
          w_se[k] = (fp.x1ok && fp.y1ok) ? __fmul_rn(fp.wx1, fp.wy1) : 0.f;
          // flags ride in the sign-free low bits of goff: keep them separate for clarity
          goff[k] = (fp.y0 * W + fp.x0) * 4 + (fp.x1ok ? 1 : 0) + (fp.y1ok ? 2 : 0);
          const int sr = fp.y0 - (th0 - HALO);
          const bool in_tile = sr >= 0 && (sr + (fp.y1ok ? 1 : 0)) < WR;
          soff[k] = in_tile ? sr * W + fp.x0 : -1;
          ooff[k] = p;
        } else {
          ooff[k] = -1; soff[k] = -1; goff[k] = 0;
          w_nw[k] = w_ne[k] = w_sw[k] = w_se[k] = 0.f;
        }
      }
    }
    const int s = it % NSTAGE;
    const uint32_t parity = (uint32_t)((it / NSTAGE) & 1);
    mbar_wait(full + s, parity);
    const T* st = reinterpret_cast<const T*>(smem_raw + (size_t)s * stage_bytes);
    const int c0 = ch * CC;
    const int rows = min(R, H - th0);
    T* out_b = out + (int64_t)tb * C * 4 * HW;

    if (tid == 0) {  // un-warped slot: shared -> global bulk store of the core rows
#pragma unroll
      for (int cc = 0; cc < CC; ++cc)
        bulk_s2g(out_b + ((int64_t)(c0 + cc) * 4 + (tt ? 3 : 0)) * HW + th0 * W, st + cc * plane + HALO * W,
                 (uint32_t)(rows * W * sizeof(T)));
      bulk_commit();
    }

    const T* gsrc = (tt ? x2 : x1) + tb * sB + (int64_t)c0 * sC;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (ooff[k] < 0) continue;
      const bool x1ok = goff[k] & 1, y1ok = goff[k] & 2;
      const int dx = x1ok ? 1 : 0;
      if (soff[k] >= 0) {
        const T* base = st + soff[k];
        const int dy = y1ok ? W : 0;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          const T* q = base + cc * plane;
          float acc = __fmul_rn(cvtf<T>(q[0]), w_nw[k]);
          if (x1ok) acc = fmaf(cvtf<T>(q[dx]), w_ne[k], acc);
          if (y1ok) acc = fmaf(cvtf<T>(q[dy]), w_sw[k], acc);
          if (x1ok && y1ok) acc = fmaf(cvtf<T>(q[dy + dx]), w_se[k], acc);
          out_b[((int64_t)(c0 + cc) * 4 + 1 + tt) * HW + ooff[k]] = fromf<T>(acc);
        }
      } else {  // footprint outside the staged rows: gather from global
        const int o = goff[k] >> 2;
        const int dy = y1ok ? W : 0;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          const T* q = gsrc + cc * sC + o;
          float acc = __fmul_rn(ldf(q), w_nw[k]);
          if (x1ok) acc = fmaf(ldf(q + dx), w_ne[k], acc);
          if (y1ok) acc = fmaf(ldf(q + dy), w_sw[k], acc);
          if (x1ok && y1ok) acc = fmaf(ldf(q + dy + dx), w_se[k], acc);
          out_b[((int64_t)(c0 + cc) * 4 + 1 + tt) * HW + ooff[k]] = fromf<T>(acc);
        }
      }
    }
    __syncthreads();  // every thread is done reading stage s
    if (tid == 0 && it + NSTAGE < total_it) {
      bulk_wait_read0();  // ... and so is the bulk store
      issue(it + NSTAGE);
    }
  }
  if (tid == 0) bulk_wait_all();
}

struct FwdPlan { int R, HALO, NP; };

template <typename T, int NP, int CC, int NSTAGE>
static int launch_fwd_tiled(const T* x1, const T* x2, int64_t sB, int64_t sC, const float* flow,
                            const float* xs, const float* ys, T* out, int B, int C, int H, int W, int R,
                            int HALO, cudaStream_t st) {
  const DeviceInfo di = device_info();
  const int nbands = (H + R - 1) / R;
  const int ntiles = 2 * B * nbands;
  const size_t smem = (size_t)NSTAGE * CC * (R + 2 * HALO) * W * sizeof(T) + NSTAGE * sizeof(uint64_t);
  if (smem > (size_t)di.smem_optin) return fail(SMOW_ERANGE, "tile does not fit shared memory (%zu B)", smem);
  auto kern = warp_fwd_tiled_kernel<T, NP, CC, NSTAGE>;
  static thread_local size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = smem;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
  if (occ < 1) occ = 1;
  const int grid = ntiles < occ * di.sms ? ntiles : occ * di.sms;
  kern<<<grid, 256, smem, st>>>(x1, x2, sB, sC, flow, xs, ys, out, C, H, W, R, HALO, nbands, ntiles);
  count_launch();
  return check_launch("warp_fwd_tiled");
}

// true when the tiled kernels can take this shape (else the caller uses the direct variant)
template <typename T>
static bool tiled_shape_ok(const void* a, const void* b, const void* c, int64_t sB, int64_t sC, int C, int H,
                           int W) {
  const int vec = 16 / (int)sizeof(T);
  return (W % vec == 0) && (sB % vec == 0) && (sC % vec == 0) && (C % 4 == 0) && W <= 1024 && aligned16(a) &&
         aligned16(b) && aligned16(c);
}

template <typename T>
static int warp_fwd_tiled(const T* x1, const T* x2, int64_t sB, int64_t sC, const float* flow,
                          const float* xs, const float* ys, T* out, int B, int C, int H, int W,
                          cudaStream_t st) {
  if (!tiled_shape_ok<T>(x1, x2, out, sB, sC, C, H, W))
    return fail(SMOW_EALIGN, "warp_fwd_tiled needs W %% %d == 0, C %% 4 == 0 and 16 B aligned tensors",
                16 / (int)sizeof(T));
  int R = option(OPT_FWD_ROWS), HALO = option(OPT_FWD_HALO);
  if (R < 1) R = 1;
  if (HALO < 1) HALO = 1;
  while (R > 1 && R * W > 1024) R >>= 1;   // at most 4 pixels per thread
  if (R > H) R = H;
  const int np = (R * W + 255) / 256;
  if (np <= 1) return launch_fwd_tiled<T, 1, 4, 3>(x1, x2, sB, sC, flow, xs, ys, out, B, C, H, W, R, HALO, st);
  if (np <= 2) return launch_fwd_tiled<T, 2, 4, 3>(x1, x2, sB, sC, flow, xs, ys, out, B, C, H, W, R, HALO, st);
  return launch_fwd_tiled<T, 4, 4, 3>(x1, x2, sB, sC, flow, xs, ys, out, B, C, H, W, R, HALO, st);
}

// ------------------------------------------------------------------------------
// backward (warp_stack_bwd_tiled.cu) and channels-last variants (warp_stack_ndhwc.cu)
// ------------------------------------------------------------------------------
template <typename T>
int warp_bwd_tiled(const T* gout, const T* x1, const T* x2, int64_t sB, int64_t sC, const float* flow,
                   const float* xs, const float* ys, T* gx1, T* gx2, float* gflow, int B, int C, int H, int W,
                   cudaStream_t st);
// channel-vectorised variant 2 (fp32 only; SMOW_ERANGE = shape not covered, caller falls back)
int warp_fwd_cvec(const float* x1, const float* x2, int64_t sB, int64_t sC, const float* flow, const float* xs,
                  const float* ys, float* out, int B, int C, int H, int W, cudaStream_t st);
int warp_bwd_cvec(const float* gout, const float* x1, const float* x2, int64_t sB, int64_t sC, const float* flow,
                  const float* xs, const float* ys, float* gx1, float* gx2, float* gflow, int B, int C, int H,
                  int W, cudaStream_t st);
int64_t warp_bwd_ndhwc_workspace_bytes(int B, int H, int W);
template <typename T>
int warp_fwd_ndhwc(const T* x1, const T* x2, int64_t sB, const float* flow, const float* xs, const float* ys,
                   T* out, int B, int C, int H, int W, cudaStream_t st);
template <typename T>
int warp_bwd_ndhwc(const T* gout, const T* x1, const T* x2, int64_t sB, const float* flow, const float* xs,
                   const float* ys, T* gx1, T* gx2, float* gflow, int B, int C, int H, int W, void* ws,
                   int64_t ws_bytes, cudaStream_t st);

}  // namespace smow
