// Hash-grid level table and per-corner arithmetic shared by the stand-alone encoder and the fused field kernels.
#pragma once
#include "common.cuh"

#define CEDNERF_MAX_LEVELS 32

struct CednerfGridLevels {
  int n_levels;
  float scale[CEDNERF_MAX_LEVELS];
  uint32_t res[CEDNERF_MAX_LEVELS];
  uint32_t size[CEDNERF_MAX_LEVELS];
  uint32_t offset[CEDNERF_MAX_LEVELS];
  uint32_t hashed[CEDNERF_MAX_LEVELS];
};

namespace {

__device__ __forceinline__ uint32_t corner_index(uint32_t gx, uint32_t gy, uint32_t gz, uint32_t res, uint32_t size,
                                                 bool hashed) {
  const uint32_t h = hashed ? (gx ^ (gy * 2654435761u) ^ (gz * 805459861u)) : (gx + gy * res + gz * res * res);
  // h % size without the integer division on the common paths: power-of-two tables (every hashed level) and
  // in-range dense cells (h < size)
  if ((size & (size - 1u)) == 0u) return h & (size - 1u);
  return h < size ? h : h % size;
}

struct Cell {
  uint32_t g[3];
  float f[3];
};

__device__ __forceinline__ Cell locate(const float* __restrict__ x, float scale) {
  Cell c;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = __fadd_rn(__fmul_rn(x[d], scale), 0.5f);
    const float fl = floorf(pos);
    c.f[d] = __fsub_rn(pos, fl);
    c.g[d] = (uint32_t)(int)fl;
  }
  return c;
}

__device__ __forceinline__ float corner_weight(const Cell& c, int corner) {
  float w = 1.f;
#pragma unroll
  for (int d = 0; d < 3; ++d) w = __fmul_rn(w, (corner >> d) & 1 ? c.f[d] : __fsub_rn(1.f, c.f[d]));
  return w;
}

// key-frame selection of the 4-D variant: ts = 3t, k = min(floor(ts), 2), tau = ts - k
// (taichi_compat: tau taken before the clamp, hash_encoder_inter.py:151-160)
__device__ __forceinline__ void keyframe(float t, int taichi_compat, int& k, float& tau) {
  const float ts = __fmul_rn(t, 3.f);
  float kf = floorf(ts);
  if (taichi_compat) {
    tau = __fsub_rn(ts, kf);
    kf = fminf(kf, 2.f);
  } else {
    kf = fminf(kf, 2.f);
    tau = __fsub_rn(ts, kf);
  }
  k = (int)kf;
}

}  // namespace
