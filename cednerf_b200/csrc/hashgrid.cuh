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

// All 8 corner weights / table indices of one cell with shared sub-expressions.  Same values as corner_weight /
// corner_index: ((1*wx)*wy)*wz == (wx*wy)*wz because 1*wx is exact, and x ^ y*P1 ^ z*P2 is formed from the two
// per-axis candidates.  Hashed levels always have a power-of-two table (size == max_params), so the modulo is a mask.
__device__ __forceinline__ void cell_weights(const float* f /*[3]*/, float* w /*[8]*/) {
  const float wx[2] = {__fsub_rn(1.f, f[0]), f[0]};
  const float wy[2] = {__fsub_rn(1.f, f[1]), f[1]};
  const float wz[2] = {__fsub_rn(1.f, f[2]), f[2]};
  float wxy[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wxy[j] = __fmul_rn(wx[j & 1], wy[j >> 1]);
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = __fmul_rn(wxy[k & 3], wz[k >> 2]);
}

__device__ __forceinline__ void cell_indices(const uint32_t* g /*[3]*/, uint32_t res, uint32_t size, uint32_t off,
                                             bool hashed, uint32_t* idx /*[8]*/) {
  if (hashed) {
    const uint32_t hx[2] = {g[0], g[0] + 1u};
    const uint32_t y0 = g[1] * 2654435761u, z0 = g[2] * 805459861u;
    const uint32_t hy[2] = {y0, y0 + 2654435761u}, hz[2] = {z0, z0 + 805459861u};
    const uint32_t mask = size - 1u;
#pragma unroll
    for (int k = 0; k < 8; ++k) idx[k] = off + ((hx[k & 1] ^ hy[(k >> 1) & 1] ^ hz[k >> 2]) & mask);
  } else {
    const uint32_t base = g[0] + g[1] * res + g[2] * res * res;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint32_t h = base + (uint32_t)(k & 1) + (uint32_t)((k >> 1) & 1) * res + (uint32_t)(k >> 2) * res * res;
      if (h >= size) h = ((size & (size - 1u)) == 0u) ? (h & (size - 1u)) : (h % size);  // out-of-range points only
      idx[k] = off + h;
    }
  }
}

__device__ __forceinline__ void cell_corners(const Cell& c, uint32_t res, uint32_t size, uint32_t off, bool hashed,
                                             uint32_t* idx /*[8]*/, float* w /*[8]*/) {
  cell_weights(c.f, w);
  cell_indices(c.g, res, size, off, hashed, idx);
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
