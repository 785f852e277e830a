// ABI bookkeeping: version, per-thread error slot, launch counter, tuning knobs.
#include "common.cuh"
#include <atomic>
#include <stdarg.h>
#include <string.h>

namespace smow {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

struct Knob { const char* key; std::atomic<int> value; };
static Knob g_knobs[OPT_COUNT] = {
    {"warp_fwd_variant", {-1}},  // -1 = auto, 0 = direct L1 gather, 1 = bulk-copy staged planes, 2 = channel-vectorised tiles
    {"warp_bwd_variant", {-1}},  // -1 = auto, 0 = atomic scatter, 1 = tiled gather, 2 = channel-vectorised gather, 3 = NDHWC gather lists (workspace), 4 = NDHWC tile gather (auto default), 0 on NDHWC = vector-atomic scatter
    {"tlerp_variant", {0}},
    {"bwd_rows", {8}},
    {"bwd_halo", {2}},
    {"fwd_rows", {8}},
    {"fwd_halo", {2}},
    {"cvec_prefetch", {3}},      // L2 bulk-prefetch distance of the channel-vectorised kernels, in chunks (0 = off)
    {"bwd_chunk_mb", {0}},       // NDHWC scatter: process the batch in chunks of about this many MB (0 = whole batch)
    {"ndhwc_bwd_rows", {0}},     // NDHWC tile gather: rows per tile (0 = auto, about 512 pixels per tile)
    {"ndhwc_bwd_pf", {-1}},      // NDHWC tile gather: next item's x taps prefetched into shared-memory slots (-1 = auto: C >= 128, 0 = off, 1 = on)
    {"tc_debug", {0}},           // bring-up switches of the tcgen05 kernels (0 in production)
    {"tok_variant", {-1}},       // tokenizer: -1 = auto (tensor-core MMA kernels for C = 16 / 32), 0 = FP32-pipe kernels
    {"bn_bwd_rows", {1}},        // fused BatchNorm + lerp backward: 1 = row-wise kernel (every gradient line fetched once), 0 = split dec / skip CTAs
};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int option(int id) { return g_knobs[id].value.load(std::memory_order_relaxed); }

DeviceInfo device_info() {
  static DeviceInfo cache[64];
  static std::atomic<int> ready[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!ready[dev].load(std::memory_order_acquire)) {
    DeviceInfo di{148, 227 * 1024};
    cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&di.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cache[dev] = di;
    ready[dev].store(1, std::memory_order_release);
  }
  return cache[dev];
}

}  // namespace smow

extern "C" {

int smow_abi_version(void) { return SMOW_ABI_VERSION; }
const char* smow_last_error(void) { return smow::g_err; }
uint64_t smow_launch_count(void) { return smow::g_launches.load(std::memory_order_relaxed); }

int smow_set_option(const char* key, int value) {
  if (!key) return SMOW_EINVAL;
  for (int i = 0; i < smow::OPT_COUNT; ++i)
    if (strcmp(key, smow::g_knobs[i].key) == 0) {
      smow::g_knobs[i].value.store(value, std::memory_order_relaxed);
      return 0;
    }
  return smow::fail(SMOW_EINVAL, "unknown option '%s'", key);
}
int smow_get_option(const char* key) {
  if (!key) return -1;
  for (int i = 0; i < smow::OPT_COUNT; ++i)
    if (strcmp(key, smow::g_knobs[i].key) == 0) return smow::option(i);
  return -1;
}

}  // extern "C"
