// Shared helpers for the cednerf_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

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

static inline int cednerf_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
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
