// Spatial bucket order of packed samples for the gradient-carrying field kernels.
//
// The visible samples of a training batch (cednerf/utils.py:74-104 after OccGridEstimator.sampling's filter) arrive
// packed per ray: ~4 per ray, the rays of a batch drawn at random over cameras, frames and pixels - so the 32 lanes of a
// warp of cednerf_field_train_fwd / _bwd sit in ~8 unrelated places of the scene, every hash-grid gather touches 32
// different sectors at every level but the coarsest, and each round of gathers waits for the slowest of ~1000 sector
// fetches (profiles/r2s_sample_order.md: the hash phase is 50 % of the training forward; with all gathers aimed at the
// same lines the kernel is 47 % faster).  Walking the same samples in the order of a 21-bit Morton key of their position
// (7 bits per axis over the field's aabb) makes neighbouring lanes neighbours in space: the coarse and middle levels
// coalesce again.  The field kernels take the order as an indirection on their EXTERNAL per-sample arrays (packed
// samples in, sigma / rgb / latent out, their gradients in); everything they keep for themselves (saved activations,
// work buffers, the table-gradient inputs) is simply laid out in the new order.
//
// A counting sort over 2^21 buckets: keys + histogram (atomics) -> exclusive scan of the buckets (1024-bin blocks, then
// their 2048 totals) -> scatter with one atomic per sample -> rank inside the bucket (ascending sample index: the order is
// a pure function of the inputs, so gradients do not change from run to run; buckets of more than 512 samples keep the
// atomics' order).  5 launches + one memset, no host read
// (`n_device`: live count on the device).
#include "common.cuh"

#define SO_BITS 7
#define SO_BINS (1 << (3 * SO_BITS))
#define SO_BLOCK 1024
#define SO_BLOCKS (SO_BINS / SO_BLOCK)
#define SO_RANK_MAX 512u

namespace {

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // bits of v (< 1024) to every third position
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__device__ __forceinline__ int64_t live_of(int64_t n, const int64_t* n_dev) {
  if (!n_dev) return n;
  const int64_t v = *n_dev;
  return v < n ? v : n;
}

__global__ void so_keys_kernel(const int64_t* __restrict__ ridx, const float* __restrict__ t0, const float* __restrict__ t1,
                               const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t n,
                               const int64_t* __restrict__ n_dev, float lo0, float lo1, float lo2, float is0, float is1,
                               float is2, uint32_t* __restrict__ keys, uint32_t* __restrict__ hist) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= live_of(n, n_dev)) return;
  const int64_t r = ridx[s];
  const float tm = 0.5f * (t0[s] + t1[s]);
  const float lo[3] = {lo0, lo1, lo2}, is[3] = {is0, is1, is2};
  uint32_t q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float u = (rays_o[3 * r + k] + rays_d[3 * r + k] * tm - lo[k]) * is[k];  // position in units of a bucket
    q[k] = (uint32_t)fminf(fmaxf(u, 0.f), (float)((1 << SO_BITS) - 1));
  }
  const uint32_t key = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
  keys[s] = key;
  atomicAdd(&hist[key], 1u);
}

// exclusive scan of each 1024-bin block in place; the block's total goes to totals[block]
__global__ void __launch_bounds__(SO_BLOCK) so_block_scan_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ totals) {
  __shared__ uint32_t ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t v = hist[blockIdx.x * SO_BLOCK + threadIdx.x];
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = ws[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    ws[lane] = t;
  }
  __syncthreads();
  hist[blockIdx.x * SO_BLOCK + threadIdx.x] = (warp ? ws[warp - 1] : 0u) + incl - v;
  if (threadIdx.x == SO_BLOCK - 1) totals[blockIdx.x] = ws[31];
}

// exclusive scan of the SO_BLOCKS (2048) block totals by one block of 1024 threads, two per thread
__global__ void __launch_bounds__(1024) so_totals_scan_kernel(uint32_t* __restrict__ totals) {
  __shared__ uint32_t ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t a = totals[2 * threadIdx.x], b = totals[2 * threadIdx.x + 1];
  uint32_t incl = a + b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = ws[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    ws[lane] = t;
  }
  __syncthreads();
  const uint32_t excl = (warp ? ws[warp - 1] : 0u) + incl - (a + b);
  totals[2 * threadIdx.x] = excl;
  totals[2 * threadIdx.x + 1] = excl + a;
}

__global__ void so_scatter_kernel(const uint32_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ n_dev,
                                  uint32_t* __restrict__ cursor, const uint32_t* __restrict__ totals,
                                  int32_t* __restrict__ unordered, uint32_t* __restrict__ slot_of) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= live_of(n, n_dev)) return;
  const uint32_t key = keys[s];
  const uint32_t p = atomicAdd(&cursor[key], 1u) + totals[key / SO_BLOCK];
  unordered[p] = (int32_t)s;
  slot_of[s] = p;
}

// The atomics above fill a bucket in arbitrary order; here every sample finds its rank among the members of its bucket
// (ascending sample index), which makes the permutation - and with it the summation order of every gradient - a pure
// function of the inputs.  After the scatter cursor[k] = start of bin k + its count = start of bin k + 1 (inside a block).
__global__ void so_rank_kernel(const uint32_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ n_dev,
                               const uint32_t* __restrict__ cursor, const uint32_t* __restrict__ totals,
                               const int32_t* __restrict__ unordered, const uint32_t* __restrict__ slot_of,
                               int32_t* __restrict__ order) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (s >= live_of(n, n_dev)) {  // the unused tail of a capacity-sized array: identity
    order[s] = (int32_t)s;
    return;
  }
  const uint32_t key = keys[s], base = totals[key / SO_BLOCK];
  const uint32_t lo = base + ((key % SO_BLOCK) ? cursor[key - 1] : 0u), hi = base + cursor[key];
  if (hi - lo > SO_RANK_MAX) {  // a crowded bucket (samples far outside the box pile up in its edge buckets): ranking is
    order[slot_of[s]] = (int32_t)s;  // quadratic in the bucket size, so such a bucket keeps the order its atomics gave it
    return;
  }
  uint32_t rank = 0;
  for (uint32_t p = lo; p < hi; ++p) rank += unordered[p] < (int32_t)s;
  order[lo + rank] = (int32_t)s;
}

}  // namespace

CEDNERF_EXPORT int64_t cednerf_sample_order_workspace_bytes(int64_t n) {
  return (int64_t)SO_BINS * 4 + SO_BLOCKS * 4 + 3 * ((n + 3) / 4 * 16);
}

// order [n] int32: a permutation of the (live) packed samples that walks them bucket by bucket of a 128^3 Morton grid
// over the box [aabb_lo, aabb_hi]; position of a sample = o + d (t0 + t1) / 2.  n_device nullable.
CEDNERF_EXPORT int cednerf_sample_order(const int64_t* ray_indices, const float* t_starts, const float* t_ends,
                                        const float* rays_o, const float* rays_d, int64_t n, const int64_t* n_device,
                                        float lo0, float lo1, float lo2, float hi0, float hi1, float hi2, void* workspace,
                                        int32_t* order, void* stream) {
  CEDNERF_REQUIRE(n >= 0 && n < (1ll << 31) && ray_indices && t_starts && t_ends && rays_o && rays_d && workspace && order,
                  "bad arguments");
  CEDNERF_REQUIRE(hi0 > lo0 && hi1 > lo1 && hi2 > lo2, "empty box");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* hist = (uint32_t*)workspace;
  uint32_t* totals = hist + SO_BINS;
  uint32_t* keys = totals + SO_BLOCKS;
  cudaMemsetAsync(hist, 0, (size_t)SO_BINS * 4, st);
  const float cells = (float)(1 << SO_BITS);
  so_keys_kernel<<<cednerf_blocks(n, 256), 256, 0, st>>>(ray_indices, t_starts, t_ends, rays_o, rays_d, n, n_device, lo0, lo1,
                                                         lo2, cells / (hi0 - lo0), cells / (hi1 - lo1), cells / (hi2 - lo2),
                                                         keys, hist);
  so_block_scan_kernel<<<SO_BLOCKS, SO_BLOCK, 0, st>>>(hist, totals);
  so_totals_scan_kernel<<<1, 1024, 0, st>>>(totals);
  int32_t* unordered = (int32_t*)(keys + (n + 3) / 4 * 4);
  uint32_t* slot_of = (uint32_t*)(unordered + (n + 3) / 4 * 4);
  so_scatter_kernel<<<cednerf_blocks(n, 256), 256, 0, st>>>(keys, n, n_device, hist, totals, unordered, slot_of);
  so_rank_kernel<<<cednerf_blocks(n, 256), 256, 0, st>>>(keys, n, n_device, hist, totals, unordered, slot_of, order);
  return cednerf_check_launch("cednerf_sample_order", 5);
}
