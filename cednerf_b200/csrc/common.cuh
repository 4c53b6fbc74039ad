// Shared helpers for the cednerf_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#define CEDNERF_EXPORT extern "C" __attribute__((visibility("default")))

// Negative library codes (positive values are cudaError_t).
#define CEDNERF_ERR_BAD_ARG (-1)
#define CEDNERF_ERR_UNSUPPORTED (-2)

void cednerf_set_error(const char* fmt, ...);
void cednerf_count_launches(int n);  // bookkeeping for cednerf_launch_count()

static inline int cednerf_check_launch(const char* what, int n_launches = 1) {
  cednerf_count_launches(n_launches);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    cednerf_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define CEDNERF_REQUIRE(cond, msg)                      \
  do {                                                  \
    if (!(cond)) {                                      \
      cednerf_set_error("%s: %s", __func__, msg);       \
      return CEDNERF_ERR_BAD_ARG;                       \
    }                                                   \
  } while (0)

// Per-device one-shot state.  Function attributes (the opt-in to > 48 KB of dynamic shared memory) and the SM count
// belong to a device, and entry points may be called from several host threads: a bit per device in an atomic word;
// losing a race only repeats an idempotent cudaFuncSetAttribute / attribute query.
#define CEDNERF_MAX_DEVICES 64
struct CednerfOncePerDevice {
  std::atomic<unsigned long long> done{0};
};

template <typename Kernel>
static inline int cednerf_opt_in_smem(Kernel kernel, int bytes, CednerfOncePerDevice& once, const char* what) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & (CEDNERF_MAX_DEVICES - 1));
  if (once.done.load(std::memory_order_acquire) & bit) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    cednerf_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  once.done.fetch_or(bit, std::memory_order_release);
  return 0;
}

static inline int cednerf_num_sms() {
  static std::atomic<int> sms[CEDNERF_MAX_DEVICES];
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<int>& slot = sms[dev & (CEDNERF_MAX_DEVICES - 1)];
  int n = slot.load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return n;
}

static inline unsigned cednerf_blocks(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  return (unsigned)(b < 1 ? 1 : b);
}

__device__ __forceinline__ float warp_incl_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// State transition at the start of a marching round of render_image_test (cednerf/utils.py:231-240), shared by
// cednerf_render_round_begin and the compaction kernel that ends the previous round.  state int32[8]: [0] alive rays of
// the round, [1] k, [2] samples per ray marched so far, [3] rays kept alive for the next round, [4] round index, [5] over.
__device__ __forceinline__ void cednerf_round_begin(int32_t* state, int64_t n_rays, int max_samples, int min_samples,
                                                    const int64_t* prev_totals, int64_t* total) {
  if (prev_totals && total) total[0] += prev_totals[0];
  int n_alive = state[3];
  state[3] = 0;
  int k = 0;
  if (n_alive <= 0 || state[2] >= max_samples) {
    n_alive = 0;
    state[5] = 1;
  } else {
    const int64_t q = n_rays / (int64_t)n_alive;  // the reference: k = max(min(n // n_alive, 64), min_samples)
    k = (int)(q < 64 ? q : 64);
    if (k < min_samples) k = min_samples;
    state[2] += k;
  }
  state[0] = n_alive;
  state[1] = k;
  state[4] += 1;
}
