// Row N4 on Blackwell tensor cores: the temporal frame mix of the decoder blocks
// (reference models/SMOW_Net.py:121-139, models/SMOW_Net_LW.py:119-137,160-175) and of the encoder's Decompose_conv
// (models/SMOW_Net.py:460-473) as a TMA-fed tcgen05 GEMM with the accumulator in tensor memory:
//
//     out[b, f, p, :] = in[b, f, p, :] @ W_0  +  in[b, (f + shift) % T, p, :] @ W_{1 + (f + own_off) % T}   (+ bias[f, :])
//
// T = 4 (decoder: cyclic exchange between the four frames) or T = 2 (encoder: T1 <-> T2 exchange).  In NDHWC memory a
// frame of one pair is a dense (pixels, C) matrix, so one CTA computes D[128 pixels, Nc] = A[128, 2C] * B[2C, Nc]:
//   * A tiles (128 pixels x 32 channels fp32 = 16 KB) arrive by TMA (cp.async.bulk.tensor, 128-byte swizzle) from the two
//     source frames; B tiles (Nc output channels x 32 input channels, K-major) from the packed weights wpack
//     [(1 + T), C_out, C_in]; a ring of up to 4 stages guarded by mbarriers;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = Nc, K = 8) per 8 channels,
//     accumulating in TMEM; tcgen05.commit frees the stage / publishes the accumulator;
//   * epilogue: 4 warps read their 32 TMEM lanes with tcgen05.ld, add the bias, stage the tile in shared memory in the
//     TMA swizzle pattern (conflict-free 16-byte stores) and one thread issues the TMA store.  The store's tensor map
//     describes a channel slice [0, C) of rows `out_pitch` floats apart, so the result can be written STRAIGHT INTO the
//     decoder's concat buffer (rows A3 + A4): the `dec` half is never copied.
// Channel counts that are not a multiple of 32 (16, 28) use the same 32-wide tiles: TMA zero-fills what lies outside
// the tensor and clips the store.  TF32 (10-bit mantissa) is the precision class the reference's own 1x1x1
// convolutions run in under torch.backends.cudnn.allow_tf32 = True; the exact-fp32 SIMT kernels of frame_mix.cu stay
// for strict-fp32 runs.
#include "common.cuh"
#include "tc05.cuh"
#include <atomic>

namespace smow {

constexpr int TC_M = 128;            // pixels per CTA tile = TMEM lanes
constexpr int TC_KC = 32;            // channels per K chunk: 128 bytes = one swizzle row
constexpr int TC_A_BYTES = TC_M * TC_KC * 4;
constexpr int TC_MAX_STAGES = 4;

// ------------------------------------------------------------------------------------------------ host: tensor maps
EncodeTiledFn tensor_map_encoder() {
  static std::atomic<EncodeTiledFn> cached{nullptr};
  EncodeTiledFn fn = cached.load(std::memory_order_acquire);
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
      q != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  cached.store(fn, std::memory_order_release);
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return fail(SMOW_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == SWIZZLE_128B_ATOM32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                              : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SMOW_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

// ------------------------------------------------------------------------------------------------ apply kernel
struct MixTcParams {
  int C, T, ntiles, shift, own_off;
  int Nc;            // MMA N = output channels per CTA (multiple of 16, <= 256)
  int kc;            // 32-channel K chunks per source frame
  int stages;        // ring depth
  int b_bytes;       // bytes of one B tile = Nc * 128
  int tmem_cols;     // power of two >= max(32, Nc)
  int out_cw;        // channels per output staging tile: 32 (128-byte swizzle) or 16 (64-byte swizzle)
  const float* bias; // [T][C] or null
  float* stats;      // persistent kernel only: per-CTA BatchNorm partial sums [cta][2][C] (sum y, sum y^2), or null
  int hw;            // pixels per frame (row mask of ragged tiles for the statistics)
};

__global__ void __launch_bounds__(128)
mix_apply_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_out, const MixTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int rest = blockIdx.x;
  const int f = rest % p.T;
  rest /= p.T;
  const int tile = rest % p.ntiles, b = rest / p.ntiles;
  const int n0 = blockIdx.y * p.Nc;
  const int stage_bytes = TC_A_BYTES + p.b_bytes;
  const int total = 2 * p.kc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer: 2 sources x kc chunks through the ring
    const int fsrc = (f + p.shift) % p.T, g = (f + p.own_off) % p.T;
    for (int it = 0; it < total; ++it) {
      const int s = it % p.stages;
      if (it >= p.stages) mbar_wait_wd(&empty_bar[s], (uint32_t)((it / p.stages) - 1) & 1u);
      const int src = it / p.kc, kk = it - src * p.kc;
      uint8_t* a_dst = ring + (size_t)s * stage_bytes;
      mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
      tma_load_3d(a_dst, &tm_in, kk * TC_KC, tile * TC_M, b * p.T + (src == 0 ? f : fsrc), &full_bar[s]);
      tma_load_2d(a_dst + TC_A_BYTES, &tm_w, kk * TC_KC, (src == 0 ? 0 : (1 + g) * p.C) + n0, &full_bar[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer: one thread, K = 8 channels per instruction
    const uint32_t idesc = umma_idesc_tf32(TC_M, p.Nc, 0, 0);
    for (int it = 0; it < total; ++it) {
      const int s = it % p.stages;
      mbar_wait_wd(&full_bar[s], (uint32_t)(it / p.stages) & 1u);
      tc_fence_after_sync();
      const int kk = it % p.kc;
      const int left = p.C - kk * TC_KC;                          // channels left in this source frame
      const int ksteps = left >= TC_KC ? TC_KC / 8 : (left + 7) / 8;
      const uint8_t* a_src = ring + (size_t)s * stage_bytes;
      const uint64_t adesc = umma_desc_kmajor(a_src, 1024, 2), bdesc = umma_desc_kmajor(a_src + TC_A_BYTES, 1024, 2);
      for (int k = 0; k < ksteps; ++k)                            // +32 bytes along K inside the swizzle atom: +2 in the address field
        umma_tf32(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, it > 0 || k > 0);
      umma_commit(&empty_bar[s]);                                 // stage free once these MMAs have read it
    }
    umma_commit(&accum_bar);                                      // accumulator complete
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers (+ bias) -> swizzled shared-memory tile -> TMA store
  mbar_wait_wd(&accum_bar, 0);
  tc_fence_after_sync();
  const int row = warp * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const float* bias = p.bias ? p.bias + (size_t)f * p.C + n0 : nullptr;
  for (int c = 0; c < p.Nc; c += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + (uint32_t)c, r);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    if (bias) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + c + j < p.C) v[j] += __ldg(bias + c + j);
    }
    uint8_t* base;
    int j0, sw;
    if (p.out_cw == 32) { base = ring + (size_t)(c >> 5) * TC_A_BYTES + row * 128; j0 = (c & 31) >> 2; sw = row & 7; }
    else                { base = ring + row * 64; j0 = 0; sw = (row >> 1) & 3; }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(base + (((j0 + q) ^ sw) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p.out_cw == 32) {
      for (int t = 0; t < p.Nc / 32; ++t)
        tma_store_3d(&tm_out, ring + (size_t)t * TC_A_BYTES, n0 + 32 * t, tile * TC_M, b * p.T + f);
    } else {
      tma_store_3d(&tm_out, ring, n0, tile * TC_M, b * p.T + f);
    }
    bulk_commit();
    bulk_wait_read0();
  }
  if (warp == 0) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ persistent apply kernel
// C <= 64 (the decoder levels that carry the bytes): the 1 + T weight matrices stay resident in shared memory, CTAs are
// persistent and warp-specialised — warp 0 feeds the A ring by TMA, warp 1 issues tcgen05.mma into one of TWO
// accumulators in tensor memory, warps 2-5 drain the other one (tcgen05.ld -> bias -> swizzled staging tile -> TMA
// store, double-buffered) — so loads, tensor-core work and the epilogue of consecutive tiles overlap.
constexpr int TCP_THREADS = 192;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __launch_bounds__(TCP_THREADS)
mix_apply_tcp_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_out, const MixTcParams p, const int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tfull_bar[2], tempty_bar[2], w_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int w_tile = p.Nc * 128;                                   // one (matrix, K chunk) weight tile
  const int w_bytes = (1 + p.T) * p.kc * w_tile;
  const int out_bytes = p.out_cw == 32 ? (p.Nc / 32) * TC_A_BYTES : TC_M * 64;
  uint8_t* wsm = base;
  uint8_t* ring = base + ((w_bytes + 1023) & ~1023);
  uint8_t* stg = ring + (size_t)p.stages * TC_A_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = 2 * p.kc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
    mbar_init(&w_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- producer: weights once, then the A tiles of every tile of this CTA
    mbar_expect_tx(&w_bar, (uint32_t)w_bytes);
    for (int m = 0; m <= p.T; ++m)
      for (int kk = 0; kk < p.kc; ++kk)
        tma_load_2d(wsm + (size_t)(m * p.kc + kk) * w_tile, &tm_w, kk * TC_KC, m * p.C, &w_bar);
    int cnt = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int f = t % p.T, rest = t / p.T, tile = rest % p.ntiles, b = rest / p.ntiles;
      const int fsrc = (f + p.shift) % p.T;
      for (int c = 0; c < chunks; ++c, ++cnt) {
        const int s = cnt % p.stages;
        if (cnt >= p.stages) mbar_wait_wd(&empty_bar[s], (uint32_t)((cnt / p.stages) - 1) & 1u);
        const int src = c / p.kc, kk = c - src * p.kc;
        mbar_expect_tx(&full_bar[s], TC_A_BYTES);
        tma_load_3d(ring + (size_t)s * TC_A_BYTES, &tm_in, kk * TC_KC, tile * TC_M, b * p.T + (src == 0 ? f : fsrc), &full_bar[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer
    const uint32_t idesc = umma_idesc_tf32(TC_M, p.Nc, 0, 0);
    mbar_wait_wd(&w_bar, 0);
    int cnt = 0, i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int f = t % p.T, g = (f + p.own_off) % p.T, acc = i & 1;
      mbar_wait_wd(&tempty_bar[acc], (uint32_t)((i >> 1) & 1) ^ 1u);       // the epilogue has drained this accumulator
      tc_fence_after_sync();
      const uint32_t d = tmem + (uint32_t)(acc * p.Nc);
      for (int c = 0; c < chunks; ++c, ++cnt) {
        const int s = cnt % p.stages;
        mbar_wait_wd(&full_bar[s], (uint32_t)(cnt / p.stages) & 1u);
        tc_fence_after_sync();
        const int src = c / p.kc, kk = c - src * p.kc;
        const int left = p.C - kk * TC_KC;
        const int ksteps = left >= TC_KC ? TC_KC / 8 : (left + 7) / 8;
        const uint64_t adesc = umma_desc_kmajor(ring + (size_t)s * TC_A_BYTES, 1024, 2);
        const uint64_t bdesc = umma_desc_kmajor(wsm + (size_t)((src == 0 ? 0 : 1 + g) * p.kc + kk) * w_tile, 1024, 2);
        for (int k = 0; k < ksteps; ++k) umma_tf32(d, adesc + 2 * k, bdesc + 2 * k, idesc, c > 0 || k > 0);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tfull_bar[acc]);
    }
  } else if (warp >= 2) {
    // ---- epilogue warps: TMEM lane quadrant = warp % 4
    const int q = warp & 3, row = q * 32 + lane;
    const bool leader = threadIdx.x == 64;
    float ssum[2] = {0.f, 0.f}, ssq[2] = {0.f, 0.f};           // BatchNorm partial sums of this thread's (row group, column)
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int f = t % p.T, rest = t / p.T, tile = rest % p.ntiles, b = rest / p.ntiles, acc = i & 1;
      uint8_t* sb = stg + (size_t)(i & 1) * out_bytes;
      mbar_wait_wd(&tfull_bar[acc], (uint32_t)(i >> 1) & 1u);
      tc_fence_after_sync();
      if (leader) bulk_wait_read1();                                        // the store that used this staging buffer is done
      epi_bar_sync();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Nc);
      const float* bias = p.bias ? p.bias + (size_t)f * p.C : nullptr;
      for (int c = 0; c < p.Nc; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c + j < p.C) v[j] += __ldg(bias + c + j);
        }
        uint8_t* dst;
        int j0, sw;
        if (p.out_cw == 32) { dst = sb + (size_t)(c >> 5) * TC_A_BYTES + row * 128; j0 = (c & 31) >> 2; sw = row & 7; }
        else                { dst = sb + row * 64; j0 = 0; sw = (row >> 1) & 3; }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          *reinterpret_cast<float4*>(dst + (((j0 + qq) ^ sw) << 4)) = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
      }
      tc_fence_before_sync();
      mbar_arrive(&tempty_bar[acc]);                                        // accumulator free for the MMA warp
      if (p.stats != nullptr) {
        // BatchNorm statistics of the block (reference models/SMOW_Net.py:136): column sums of the staged tile, read back
        // conflict-free through the swizzle; rows past the frame's last pixel (ragged tile) are masked out
        epi_bar_sync();
        const int et = threadIdx.x - 64, valid = p.hw - tile * TC_M < TC_M ? p.hw - tile * TC_M : TC_M;
        if (p.out_cw == 32) {
          const int c = et & 31, rg = et >> 5;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k >= p.Nc / 32) break;
            const uint8_t* tb = sb + (size_t)k * TC_A_BYTES + (c & 3) * 4;
            float a = 0.f, b2 = 0.f;
            for (int r = rg * 32; r < rg * 32 + 32; ++r) {
              const float v = r < valid ? *reinterpret_cast<const float*>(tb + r * 128 + ((((c >> 2) ^ (r & 7))) << 4)) : 0.f;
              a += v; b2 = fmaf(v, v, b2);
            }
            ssum[k] += a; ssq[k] += b2;
          }
        } else {
          const int c = et & 15, rg = et >> 4;
          float a = 0.f, b2 = 0.f;
          for (int r = rg * 16; r < rg * 16 + 16; ++r) {
            const float v = r < valid ? *reinterpret_cast<const float*>(sb + r * 64 + ((((c >> 2) ^ ((r >> 1) & 3))) << 4) + (c & 3) * 4) : 0.f;
            a += v; b2 = fmaf(v, v, b2);
          }
          ssum[0] += a; ssq[0] += b2;
        }
      }
      fence_proxy_async();
      epi_bar_sync();
      if (leader) {
        if (p.out_cw == 32) {
          for (int k = 0; k < p.Nc / 32; ++k) tma_store_3d(&tm_out, sb + (size_t)k * TC_A_BYTES, 32 * k, tile * TC_M, b * p.T + f);
        } else {
          tma_store_3d(&tm_out, sb, 0, tile * TC_M, b * p.T + f);
        }
        bulk_commit();
      }
    }
    if (leader) bulk_wait_read0();
    if (p.stats != nullptr) {
      // row groups -> one (sum, sum of squares) per channel and CTA, added in a fixed order
      epi_bar_sync();
      float* sc = reinterpret_cast<float*>(stg);                           // [groups][2][Nc]
      const int et = threadIdx.x - 64;
      const int groups = p.out_cw == 32 ? 4 : 8, cw = p.out_cw, c = et % cw, rg = et / cw;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k >= (p.out_cw == 32 ? p.Nc / 32 : 1)) break;
        sc[(rg * 2 + 0) * p.Nc + k * 32 + c] = ssum[k];
        sc[(rg * 2 + 1) * p.Nc + k * 32 + c] = ssq[k];
      }
      epi_bar_sync();
      for (int e = et; e < 2 * p.Nc; e += 128) {
        const int which = e / p.Nc, cc = e - which * p.Nc;
        if (cc < p.C) {
          float t = 0.f;
          for (int g2 = 0; g2 < groups; ++g2) t += sc[(g2 * 2 + which) * p.Nc + cc];
          p.stats[((size_t)blockIdx.x * 2 + which) * p.C + cc] = t;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

// tile plan for C channels; false when the channel count is not covered
static bool tc_plan(int C, MixTcParams& p, int& nsplit) {
  if (C < 16 || C > 512 || C % 4 != 0) return false;
  const int cpad = (C + 15) / 16 * 16;                 // 28 -> 32
  nsplit = cpad > 256 ? 2 : 1;                         // 320 -> 2 x 160, 512 -> 2 x 256
  if (cpad % (16 * nsplit) != 0) return false;
  p.Nc = cpad / nsplit;
  p.out_cw = (p.Nc % 32 == 0) ? 32 : 16;
  if (p.out_cw == 16 && p.Nc != 16) return false;      // 16, 28, and every multiple of 32 (x2 above 256)
  p.kc = (C + TC_KC - 1) / TC_KC;
  p.b_bytes = p.Nc * 128;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.Nc) p.tmem_cols *= 2;
  const int stage_bytes = TC_A_BYTES + p.b_bytes;
  p.stages = 2 * p.kc < TC_MAX_STAGES ? 2 * p.kc : TC_MAX_STAGES;
  while (p.stages > 2 && (size_t)p.stages * stage_bytes + 2048 > 200 * 1024) --p.stages;
  return true;
}

}  // namespace smow

using namespace smow;

extern "C" {

int smow_frame_mix_tc_supported(int C, int T) {
  MixTcParams p;
  int nsplit;
  return (tc_plan(C, p, nsplit) && (T == 2 || T == 4)) ? 1 : 0;
}

static int frame_mix_apply_tc_impl(const float* in, const float* wpack, const float* bias, float* out, int B, int C, int T,
                                   int64_t hw, int64_t out_pitch, int shift, int own_off, float* stats, int* stats_ctas,
                                   void* stream);

int smow_frame_mix_apply_tc(const float* in, const float* wpack, const float* bias, float* out, int B, int C, int T,
                            int64_t hw, int64_t out_pitch, int shift, int own_off, void* stream) {
  return frame_mix_apply_tc_impl(in, wpack, bias, out, B, C, T, hw, out_pitch, shift, own_off, nullptr, nullptr, stream);
}

// number of per-CTA partial rows smow_frame_mix_apply_tc_stats writes (0: this shape has no statistics epilogue)
int smow_frame_mix_stats_parts(int B, int C, int T, int64_t hw) {
  MixTcParams p;
  int nsplit = 1;
  if (B <= 0 || hw <= 0 || (T != 2 && T != 4) || !tc_plan(C, p, nsplit) || p.kc > 2 || nsplit != 1) return 0;
  const int w_bytes = ((1 + T) * p.kc * p.Nc * 128 + 1023) & ~1023;
  const int out_b = p.out_cw == 32 ? (p.Nc / 32) * TC_A_BYTES : TC_M * 64;
  const int fixed = w_bytes + 2 * out_b + 1024;
  const int per_sm = (110 * 1024 - fixed) / TC_A_BYTES >= 3 ? 2 : 1;
  const int64_t total = (int64_t)B * T * ((hw + TC_M - 1) / TC_M);
  const int64_t nctas = (int64_t)device_info().sms * per_sm;
  return (int)(nctas < total ? nctas : total);
}

int smow_frame_mix_apply_tc_stats(const float* in, const float* wpack, const float* bias, float* out, float* stats, int B,
                                  int C, int T, int64_t hw, int64_t out_pitch, int shift, int own_off, void* stream) {
  if (!stats || !aligned16(stats)) return fail(SMOW_EINVAL, "frame_mix_tc: statistics buffer missing / misaligned");
  int ctas = 0;
  const int rc = frame_mix_apply_tc_impl(in, wpack, bias, out, B, C, T, hw, out_pitch, shift, own_off, stats, &ctas, stream);
  if (rc == 0 && ctas != smow_frame_mix_stats_parts(B, C, T, hw)) return fail(SMOW_EINVAL, "frame_mix_tc: statistics plan mismatch");
  return rc;
}

static int frame_mix_apply_tc_impl(const float* in, const float* wpack, const float* bias, float* out, int B, int C, int T,
                                   int64_t hw, int64_t out_pitch, int shift, int own_off, float* stats, int* stats_ctas,
                                   void* stream) {
  if (!in || !wpack || !out || B <= 0 || hw <= 0) return fail(SMOW_EINVAL, "frame_mix_tc: bad shape / null pointer");
  if (!smow_frame_mix_tc_supported(C, T)) return fail(SMOW_EDTYPE, "frame_mix_tc: unsupported C = %d / T = %d", C, T);
  if (out_pitch < C || out_pitch % 4 != 0) return fail(SMOW_EINVAL, "frame_mix_tc: out_pitch must be >= C and a multiple of 4");
  if (!aligned16(in) || !aligned16(out) || !aligned16(wpack)) return fail(SMOW_EALIGN, "frame_mix_tc: 16 B alignment");
  const int64_t ntiles = (hw + TC_M - 1) / TC_M;
  if ((int64_t)B * T * ntiles > 0x7fffffffll || hw > 0x7fffffffll) return fail(SMOW_ERANGE, "frame_mix_tc: tensor too large");
  MixTcParams p;
  int nsplit = 1;
  if (!tc_plan(C, p, nsplit)) return fail(SMOW_EDTYPE, "frame_mix_tc: unsupported C = %d", C);
  p.C = C; p.T = T; p.ntiles = (int)ntiles; p.shift = ((shift % T) + T) % T; p.own_off = ((own_off % T) + T) % T;
  p.bias = bias;
  p.stats = nullptr;
  p.hw = (int)hw;
  const int stage_bytes = TC_A_BYTES + p.b_bytes;
  size_t smem = (size_t)p.stages * stage_bytes;
  const size_t out_bytes = p.out_cw == 32 ? (size_t)(p.Nc / 32) * TC_A_BYTES : (size_t)TC_M * 64;
  if (smem < out_bytes) smem = out_bytes;
  smem += 1024;                                        // alignment slack for the 1024-byte swizzle atoms

  CUtensorMap tm_in, tm_w, tm_out;
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)hw, (uint64_t)B * T};
    const uint64_t str[2] = {(uint64_t)C * 4, (uint64_t)hw * C * 4};
    const uint32_t box[3] = {TC_KC, TC_M, 1};
    if (int rc = make_tensor_map(&tm_in, in, 3, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)(1 + T) * C};
    const uint64_t str[1] = {(uint64_t)C * 4};
    const uint32_t box[2] = {TC_KC, (uint32_t)p.Nc};
    if (int rc = make_tensor_map(&tm_w, wpack, 2, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)hw, (uint64_t)B * T};
    const uint64_t str[2] = {(uint64_t)out_pitch * 4, (uint64_t)hw * out_pitch * 4};
    const uint32_t box[3] = {(uint32_t)p.out_cw, TC_M, 1};
    if (int rc = make_tensor_map(&tm_out, out, 3, dims, str, box, p.out_cw * 4)) return rc;
  }
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(mix_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("frame_mix_tc (shared-memory opt-in)");
  const int64_t total_tiles = (int64_t)B * T * ntiles;
  if (p.kc <= 2 && nsplit == 1 && !(option(OPT_TC_DEBUG) & 32)) {
    // persistent, warp-specialised variant with resident weights
    const int w_bytes = ((1 + T) * p.kc * p.Nc * 128 + 1023) & ~1023;
    const int out_b = p.out_cw == 32 ? (p.Nc / 32) * TC_A_BYTES : TC_M * 64;
    const int fixed = w_bytes + 2 * out_b + 1024;
    int stages = (110 * 1024 - fixed) / TC_A_BYTES;                 // two CTAs per SM when the weights allow it
    int per_sm = 2;
    if (stages < 3) { stages = (220 * 1024 - fixed) / TC_A_BYTES; per_sm = 1; }
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(SMOW_EDTYPE, "frame_mix_tc: shared memory plan failed for C = %d", C);
    MixTcParams pp = p;
    pp.stats = stats;
    pp.stages = stages;
    pp.tmem_cols = 32;
    while (pp.tmem_cols < 2 * p.Nc) pp.tmem_cols *= 2;
    const size_t psmem = (size_t)fixed + (size_t)stages * TC_A_BYTES;
    if (psmem > 48 * 1024 &&
        cudaFuncSetAttribute(mix_apply_tcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem) != cudaSuccess)
      return check_launch("frame_mix_tc (shared-memory opt-in)");
    int64_t nctas = (int64_t)device_info().sms * per_sm;
    if (nctas > total_tiles) nctas = total_tiles;
    mix_apply_tcp_kernel<<<(unsigned)nctas, TCP_THREADS, psmem, (cudaStream_t)stream>>>(tm_in, tm_w, tm_out, pp, (int)total_tiles);
    if (stats_ctas) *stats_ctas = (int)nctas;
    count_launch();
    return check_launch("frame_mix_apply_tc");
  }
  if (stats) return fail(SMOW_EDTYPE, "frame_mix_tc: the statistics epilogue exists for C <= 64 only (got C = %d)", C);
  const dim3 grid((unsigned)total_tiles, (unsigned)nsplit);
  mix_apply_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tm_in, tm_w, tm_out, p);
  count_launch();
  return check_launch("frame_mix_apply_tc");
}

}  // extern "C"

// =====================================================================================================================
// Weight gradients on the tensor cores.  For every (pair b, 64-pixel tile, frame f):
//     D_0        += X_f^T            G_f          (-> dW_0, the shared matrix)
//     D_{1+g(f)} += X_{(f+shift)%T}^T G_f          (-> dW_{1+g}, g = (f + own_off) % T)
// with M = input channels, N = output channels, K = pixels.  The NDHWC tiles are exactly what TMA delivers
// ([pixels][32 channels], 128-byte swizzle), and with the CHANNEL index contiguous they are "MN-major" operands for
// both A and B: no transposition anywhere.  A CTA owns a range of (pair, tile) units, keeps its (1 + T) accumulators
// [128 x Ncw] in tensor memory for the whole range and writes ONE partial per CTA; a second kernel adds the partials in
// index order (deterministic).  M is always 128: for C < 128 the upper accumulator rows hold products of whatever
// follows the tile in shared memory and are never read.
// =====================================================================================================================
namespace smow {

constexpr int WG_PX = 64;                         // pixels per tile (K per stage; 8 MMAs of K = 8)
constexpr int WG_BLK_BYTES = WG_PX * 128;         // one [64 px][32 ch] block = 8 KB
constexpr int WG_MAX_STAGES = 4;

struct WgradTcParams {
  int C, T, shift, own_off;
  int ntiles;          // 64-pixel tiles per frame
  int nunits;          // B * ntiles
  int units_per_cta;
  int mblk;            // 32-channel blocks of the A operands loaded per step (<= 4)
  int nblk;            // 32-channel blocks of the B operand = Ncw / 32 (Ncw = 32 * nblk, or 16 for C = 16)
  int Ncw;             // MMA N
  int stages, stage_bytes, tmem_cols;
  int dbg;
  float* part;         // [ctas][mchunks*nchunks ...] see below
};

// partial layout: part[((cta * (1+T) + slot) * C + ci) * C + co]
__global__ void __launch_bounds__(128)
mix_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g, const WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.z * p.Ncw;      // channel offsets of this CTA's [128 x Ncw] block
  const int u0 = blockIdx.x * p.units_per_cta;
  int u1 = u0 + p.units_per_cta;
  if (u1 > p.nunits) u1 = p.nunits;
  const int steps = (u1 - u0) * p.T;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int a_bytes = p.mblk * WG_BLK_BYTES;                     // one A operand (all its channel blocks)

  if (warp == 0 && lane == 0) {
    for (int it = 0; it < steps; ++it) {
      const int s = it % p.stages;
      if (it >= p.stages) mbar_wait_wd(&empty_bar[s], (uint32_t)((it / p.stages) - 1) & 1u);
      const int u = u0 + it / p.T, f = it % p.T;
      const int b = u / p.ntiles, tile = u - b * p.ntiles;
      const int fs = (f + p.shift) % p.T;
      uint8_t* dst = ring + (size_t)s * p.stage_bytes;
      mbar_expect_tx(&full_bar[s], (uint32_t)((2 * p.mblk + p.nblk) * WG_BLK_BYTES));
      for (int j = 0; j < p.mblk; ++j) {
        tma_load_3d(dst + j * WG_BLK_BYTES, &tm_x, m0 + 32 * j, tile * WG_PX, b * p.T + f, &full_bar[s]);
        tma_load_3d(dst + a_bytes + j * WG_BLK_BYTES, &tm_x, m0 + 32 * j, tile * WG_PX, b * p.T + fs, &full_bar[s]);
      }
      for (int j = 0; j < p.nblk; ++j)
        tma_load_3d(dst + 2 * a_bytes + j * WG_BLK_BYTES, &tm_g, n0 + 32 * j, tile * WG_PX, b * p.T + f, &full_bar[s]);
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, p.Ncw, 1, 1);
    uint32_t started = 0;                                         // bit m set once accumulator m holds data
    for (int it = 0; it < steps; ++it) {
      const int s = it % p.stages;
      mbar_wait_wd(&full_bar[s], (uint32_t)(it / p.stages) & 1u);
      tc_fence_after_sync();
      const int f = it % p.T, g = (f + p.own_off) % p.T;
      const uint8_t* src = ring + (size_t)s * p.stage_bytes;
      // MN-major TF32 operands exist in ONE shared-memory layout: 128-byte rows swizzled in 32-byte units (layout type 1,
      // TMA mode 128B_ATOM_32B); its K atom is 4 pixel rows = 512 bytes (SBO), the next 32-channel block is LBO away
      uint32_t lbo = WG_BLK_BYTES, sbo = (p.dbg & 16) ? 1024 : 512;
      if (p.dbg & 1) { const uint32_t t = lbo; lbo = sbo; sbo = t; }
      const uint64_t a1 = umma_desc_mnmajor(src, lbo, sbo, 1);
      const uint64_t a2 = umma_desc_mnmajor(src + a_bytes, lbo, sbo, 1);
      const uint64_t bd = umma_desc_mnmajor(src + 2 * a_bytes, lbo, sbo, 1);
      const uint32_t d0 = tmem, dg = tmem + (uint32_t)((1 + g) * p.Ncw);
      for (int k = 0; k < WG_PX / 8; ++k) {                       // 8 pixels = one 1024-byte swizzle atom per step: +64
        umma_tf32(d0, a1 + 64 * k, bd + 64 * k, idesc, (started & 1u) || k > 0);
        umma_tf32(dg, a2 + 64 * k, bd + 64 * k, idesc, ((started >> (1 + g)) & 1u) || k > 0);
      }
      started |= 1u | (1u << (1 + g));
      umma_commit(&empty_bar[s]);
    }
    umma_commit(&accum_bar);
  }
  __syncwarp();

  mbar_wait_wd(&accum_bar, 0);
  tc_fence_after_sync();
  const int ci = m0 + warp * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const bool all = steps >= p.T;                                   // every accumulator was written at least once
  for (int m = 0; m <= p.T; ++m) {
    float* dst = p.part + (((size_t)blockIdx.x * (1 + p.T) + m) * p.C + ci) * p.C + n0;
    for (int c = 0; c < p.Ncw; c += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + (uint32_t)(m * p.Ncw + c), r);
      tmem_ld_wait();
      if (ci < p.C) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (n0 + c + 4 * q < p.C)
            *reinterpret_cast<float4*>(dst + c + 4 * q) =
                all ? make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                  __uint_as_float(r[4 * q + 3]))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

// gw[m][ci][co] = sum over CTAs of part[cta][m][ci][co].  grid (ceil(C*C/32), 1+T), block 256 = 32 elements x 8 CTA slices:
// every thread adds every 8th partial (independent loads in flight), the 8 slices are then added in a fixed order.
__global__ void __launch_bounds__(256)
mix_wgrad_tc_combine_kernel(const float* __restrict__ part, float* __restrict__ gw, int CC, int nm, int nctas) {
  __shared__ float red[8][32];
  const int el = threadIdx.x & 31, sl = threadIdx.x >> 5, e = blockIdx.x * 32 + el, m = blockIdx.y;
  float t0 = 0.f, t1 = 0.f;
  if (e < CC) {
    const float* src = part + (size_t)m * CC + e;
    int c = sl;
    for (; c + 8 < nctas; c += 16) { t0 += src[(size_t)c * nm * CC]; t1 += src[(size_t)(c + 8) * nm * CC]; }
    if (c < nctas) t0 += src[(size_t)c * nm * CC];
  }
  red[sl][el] = t0 + t1;
  __syncthreads();
  if (sl == 0 && e < CC) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += red[k][el];
    gw[(size_t)m * CC + e] = r;
  }
}

// ---- T = 4, C <= 32 (the big decoder levels): the four frames of a 64-pixel tile are stacked along M ---------------------
// A = [X_0 ; X_1 ; X_2 ; X_3] (4 x 32 channel rows = the full M = 128; the four tiles sit 8 KB apart, which is exactly the
// descriptor's MN-block stride), B = G_f:  D_f[f'*32 + ci, co] = sum_p X_f'[p, ci] G_f[p, co]  for all f' in ONE instruction
// stream per f.  Needed blocks: f' = f (-> dW_0) and f' = (f + shift) % 4 (-> dW_{1+g}); every tile is loaded exactly once.
constexpr int WG4_STAGE_BYTES = 8 * WG_BLK_BYTES;        // X_0..3, G_0..3
constexpr int WG4_STAGES = 3;

struct Wgrad4Params {
  int C, shift, own_off, ntiles, nunits, Ncw;
  float* part;                                           // [cta][8 slots][C][C]: slots 0..3 dW_0 from frame w, 4..7 dW_{1+g}
};

__global__ void __launch_bounds__(128)
mix_wgrad_tc4_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g, const Wgrad4Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG4_STAGES], empty_bar[WG4_STAGES], accum_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tmem_cols = 4 * p.Ncw < 32 ? 32 : 4 * p.Ncw;            // 64 or 128
  int nmine = 0;
  for (int u = blockIdx.x; u < p.nunits; u += gridDim.x) ++nmine;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG4_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0 && lane == 0) {
    int it = 0;
    for (int u = blockIdx.x; u < p.nunits; u += gridDim.x, ++it) {
      const int s = it % WG4_STAGES;
      if (it >= WG4_STAGES) mbar_wait_wd(&empty_bar[s], (uint32_t)((it / WG4_STAGES) - 1) & 1u);
      const int b = u / p.ntiles, tile = u - b * p.ntiles;
      uint8_t* dst = ring + (size_t)s * WG4_STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], WG4_STAGE_BYTES);
      for (int f = 0; f < 4; ++f) {
        tma_load_3d(dst + f * WG_BLK_BYTES, &tm_x, 0, tile * WG_PX, b * 4 + f, &full_bar[s]);
        tma_load_3d(dst + (4 + f) * WG_BLK_BYTES, &tm_g, 0, tile * WG_PX, b * 4 + f, &full_bar[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, p.Ncw, 1, 1);
    for (int it = 0; it < nmine; ++it) {
      const int s = it % WG4_STAGES;
      mbar_wait_wd(&full_bar[s], (uint32_t)(it / WG4_STAGES) & 1u);
      tc_fence_after_sync();
      const uint8_t* src = ring + (size_t)s * WG4_STAGE_BYTES;
      const uint64_t ad = umma_desc_mnmajor(src, WG_BLK_BYTES, 512, 1);
      for (int f = 0; f < 4; ++f) {
        const uint64_t bd = umma_desc_mnmajor(src + (4 + f) * WG_BLK_BYTES, WG_BLK_BYTES, 512, 1);
        const uint32_t d = tmem + (uint32_t)(f * p.Ncw);
        for (int k = 0; k < WG_PX / 8; ++k) umma_tf32(d, ad + 64 * k, bd + 64 * k, idesc, it > 0 || k > 0);
      }
      umma_commit(&empty_bar[s]);
    }
    umma_commit(&accum_bar);
  }
  __syncwarp();

  mbar_wait_wd(&accum_bar, 0);
  tc_fence_after_sync();
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const int ci = lane;
  const int f_own = (warp - p.shift + 4) & 3, g = (f_own + p.own_off) & 3;
  for (int half = 0; half < 2; ++half) {
    const int fsel = half == 0 ? warp : f_own, slot = half == 0 ? warp : 4 + g;
    float* dst = p.part + (((size_t)blockIdx.x * 8 + slot) * p.C + ci) * p.C;
    for (int c = 0; c < p.Ncw; c += 16) {
      uint32_t r[16];
      tmem_ld16(taddr + (uint32_t)(fsel * p.Ncw + c), r);
      tmem_ld_wait();
      if (ci < p.C) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c + 4 * q < p.C)
            *reinterpret_cast<float4*>(dst + c + 4 * q) =
                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                            __uint_as_float(r[4 * q + 3]));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, (uint32_t)tmem_cols);
  }
}

// gw[0] = sum over CTAs and slots 0..3; gw[1+g] = sum over CTAs of slot 4+g.  grid (ceil(C*C/32), 5), block 256 = 32 elements
// x 8 CTA slices, fixed-order slice reduction.
__global__ void __launch_bounds__(256)
mix_wgrad_tc4_combine_kernel(const float* __restrict__ part, float* __restrict__ gw, int CC, int nctas) {
  __shared__ float red[8][32];
  const int el = threadIdx.x & 31, sl = threadIdx.x >> 5, e = blockIdx.x * 32 + el, m = blockIdx.y;
  float t0 = 0.f, t1 = 0.f;
  if (e < CC) {
    if (m == 0) {
      for (int c = sl; c < nctas; c += 8) {
        const float* src = part + (size_t)c * 8 * CC + e;
        t0 += src[0] + src[(size_t)CC];
        t1 += src[(size_t)2 * CC] + src[(size_t)3 * CC];
      }
    } else {
      const float* src = part + (size_t)(3 + m) * CC + e;
      int c = sl;
      for (; c + 8 < nctas; c += 16) { t0 += src[(size_t)c * 8 * CC]; t1 += src[(size_t)(c + 8) * 8 * CC]; }
      if (c < nctas) t0 += src[(size_t)c * 8 * CC];
    }
  }
  red[sl][el] = t0 + t1;
  __syncthreads();
  if (sl == 0 && e < CC) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += red[k][el];
    gw[(size_t)m * CC + e] = r;
  }
}

static bool wgrad_tc4_ok(int C, int T) { return T == 4 && C >= 16 && C <= 32 && C % 4 == 0 && !(option(OPT_TC_DEBUG) & 64); }
static int wgrad_tc4_ctas(int B, int64_t hw) {
  const int64_t nunits = (int64_t)B * ((hw + WG_PX - 1) / WG_PX);
  const int sms = device_info().sms;
  return (int)(nunits < sms ? nunits : sms);
}

static bool wgrad_tc_plan(int B, int C, int T, int64_t hw, WgradTcParams& p, int& nctas, int& mchunks, int& nchunks) {
  if (C < 16 || C > 512 || C % 4 != 0 || (T != 2 && T != 4)) return false;
  if (C % 32 != 0 && C > 32) return false;                        // 16, 20, 24, 28, 32, then multiples of 32
  p.C = C; p.T = T;
  const int cblocks = (C + 31) / 32;
  p.mblk = cblocks < 4 ? cblocks : 4;
  mchunks = (cblocks + 3) / 4;
  // accumulators: (1 + T) * Ncw columns <= 512
  int ncw = 32 * cblocks;
  const int cap = (512 / (1 + T)) / 32 * 32;                      // T = 4: 96 ; T = 2: 160
  if (ncw > cap) ncw = cap;
  while ((32 * cblocks) % ncw != 0) ncw -= 32;
  if (C <= 16) ncw = 16;
  p.Ncw = ncw;
  p.nblk = (ncw + 31) / 32;
  nchunks = C <= 16 ? 1 : (32 * cblocks) / ncw;
  p.tmem_cols = 32;
  while (p.tmem_cols < (1 + T) * ncw) p.tmem_cols *= 2;
  p.ntiles = (int)((hw + WG_PX - 1) / WG_PX);
  p.nunits = B * p.ntiles;
  int target = 2 * device_info().sms;
  // the per-CTA partials must stay small next to the tensors themselves: <= 1/8 of x + gy, or 16 MB for small tensors
  int64_t budget = ((int64_t)2 * B * T * hw * C * 4) / 8;
  if (budget < (16ll << 20)) budget = 16ll << 20;
  const int64_t cap_ctas = budget / ((int64_t)(1 + T) * C * C * 4);
  if (cap_ctas < target) target = cap_ctas < 1 ? 1 : (int)cap_ctas;
  nctas = p.nunits < target ? p.nunits : target;
  p.units_per_cta = (p.nunits + nctas - 1) / nctas;
  nctas = (p.nunits + p.units_per_cta - 1) / p.units_per_cta;
  p.stage_bytes = (2 * p.mblk + p.nblk) * WG_BLK_BYTES;
  p.stages = WG_MAX_STAGES;
  while (p.stages > 2 && (size_t)p.stages * p.stage_bytes > 150 * 1024) --p.stages;
  return true;
}

}  // namespace smow

extern "C" {

int64_t smow_frame_mix_wgrad_tc_workspace_bytes(int B, int C, int T, int64_t hw) {
  WgradTcParams p;
  int nctas, mch, nch;
  if (B <= 0 || hw <= 0) return 0;
  if (wgrad_tc4_ok(C, T)) return (int64_t)wgrad_tc4_ctas(B, hw) * 8 * C * C * (int64_t)sizeof(float);
  if (!wgrad_tc_plan(B, C, T, hw, p, nctas, mch, nch)) return 0;
  return (int64_t)nctas * (1 + T) * C * C * (int64_t)sizeof(float);
}

int smow_frame_mix_wgrad_tc(const float* x, const float* gy, float* gw, int B, int C, int T, int64_t hw, int shift,
                            int own_off, void* ws, int64_t ws_bytes, void* stream) {
  if (!x || !gy || !gw || B <= 0 || hw <= 0) return fail(SMOW_EINVAL, "frame_mix_wgrad_tc: bad shape / null pointer");
  if (wgrad_tc4_ok(C, T)) {
    if (!aligned16(x) || !aligned16(gy) || !aligned16(gw) || !ws || !aligned16(ws) ||
        ws_bytes < smow_frame_mix_wgrad_tc_workspace_bytes(B, C, T, hw))
      return fail(SMOW_EINVAL, "frame_mix_wgrad_tc: workspace of smow_frame_mix_wgrad_tc_workspace_bytes() bytes required");
    Wgrad4Params q;
    q.C = C; q.shift = ((shift % 4) + 4) % 4; q.own_off = ((own_off % 4) + 4) % 4;
    q.ntiles = (int)((hw + WG_PX - 1) / WG_PX); q.nunits = B * q.ntiles; q.Ncw = C <= 16 ? 16 : 32;
    q.part = reinterpret_cast<float*>(ws);
    CUtensorMap tm_x, tm_g;
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)hw, (uint64_t)B * 4};
    const uint64_t str[2] = {(uint64_t)C * 4, (uint64_t)hw * C * 4};
    const uint32_t box[3] = {32, WG_PX, 1};
    if (int rc = make_tensor_map(&tm_x, x, 3, dims, str, box, SWIZZLE_128B_ATOM32)) return rc;
    if (int rc = make_tensor_map(&tm_g, gy, 3, dims, str, box, SWIZZLE_128B_ATOM32)) return rc;
    const size_t smem4 = (size_t)WG4_STAGES * WG4_STAGE_BYTES + 1024;
    if (cudaFuncSetAttribute(mix_wgrad_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4) != cudaSuccess)
      return check_launch("frame_mix_wgrad_tc (shared-memory opt-in)");
    const int nctas = wgrad_tc4_ctas(B, hw);
    cudaStream_t st4 = (cudaStream_t)stream;
    mix_wgrad_tc4_kernel<<<nctas, 128, smem4, st4>>>(tm_x, tm_g, q);
    mix_wgrad_tc4_combine_kernel<<<dim3((C * C + 31) / 32, 5), 256, 0, st4>>>(q.part, gw, C * C, nctas);
    count_launch(2);
    return check_launch("frame_mix_wgrad_tc");
  }
  WgradTcParams p;
  int nctas, mch, nch;
  if (!wgrad_tc_plan(B, C, T, hw, p, nctas, mch, nch)) return fail(SMOW_EDTYPE, "frame_mix_wgrad_tc: unsupported C = %d / T = %d", C, T);
  if (!aligned16(x) || !aligned16(gy) || !aligned16(gw) || !ws || !aligned16(ws) ||
      ws_bytes < smow_frame_mix_wgrad_tc_workspace_bytes(B, C, T, hw))
    return fail(SMOW_EINVAL, "frame_mix_wgrad_tc: workspace of smow_frame_mix_wgrad_tc_workspace_bytes() bytes required");
  p.shift = ((shift % T) + T) % T; p.own_off = ((own_off % T) + T) % T;
  p.part = reinterpret_cast<float*>(ws);
  p.dbg = option(OPT_TC_DEBUG);
  CUtensorMap tm_x, tm_g;
  const uint64_t dims[3] = {(uint64_t)C, (uint64_t)hw, (uint64_t)B * T};
  const uint64_t str[2] = {(uint64_t)C * 4, (uint64_t)hw * C * 4};
  const uint32_t box[3] = {32, WG_PX, 1};
  if (int rc = make_tensor_map(&tm_x, x, 3, dims, str, box, SWIZZLE_128B_ATOM32)) return rc;
  if (int rc = make_tensor_map(&tm_g, gy, 3, dims, str, box, SWIZZLE_128B_ATOM32)) return rc;
  // M = 128 reads four 32-channel blocks from the A tile's base: pad the ring so that the read stays inside the allocation
  const size_t smem = (size_t)p.stages * p.stage_bytes + 4 * WG_BLK_BYTES + 1024;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(mix_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return check_launch("frame_mix_wgrad_tc (shared-memory opt-in)");
  cudaStream_t st = (cudaStream_t)stream;
  mix_wgrad_tc_kernel<<<dim3(nctas, mch, nch), 128, smem, st>>>(tm_x, tm_g, p);
  mix_wgrad_tc_combine_kernel<<<dim3((C * C + 31) / 32, 1 + T), 256, 0, st>>>(p.part, gw, C * C, 1 + T, nctas);
  count_launch(2);
  return check_launch("frame_mix_wgrad_tc");
}

}  // extern "C"
