// Importance-sampled training batches: the training branch of SubjectLoader.fetch_data,
// datasets/dnerf_3d_video_IS.py:401-445 (ISG / IST weights) and :447-497 (pixel gather + ray generation).
//
// The reference draws `batch_size = num_rays / s^2` cells of the s-times-subsampled weight map with
// torch.multinomial(weights[subset], batch_size) (replacement = False) after thinning the 1.85 G weights to a uniform
// random subset of sampling_batch_size entries, expands every drawn cell into its s x s pixels, gathers their colours
// and generates their rays.  torch.multinomial without replacement IS a top-k: q = weights / Exp(1) noise, the
// batch_size largest q win (aten/src/ATen/native/Sampling / multinomial: exponential_(1), div, topk - the "exponential
// race" form of Efraimidis-Spirakis sampling).  Here:
//   * cednerf_importance_keys : q as order-preserving 32-bit keys (non-negative floats order like their bit patterns),
//   * cednerf_topk_select     : the k largest keys by a three-pass radix select (11 + 11 + 10 bits: histogram of the
//                               candidates' next digit, pick the digit in which the k-th largest falls) and an ORDERED
//                               compaction (two-counter block scan: keys above the threshold, keys equal to it - ties at
//                               the threshold go to the lowest positions, so the result is deterministic),
//   * cednerf_importance_batch: cell -> s x s pixels -> colour / 255, ray, timestamp, image id in one launch (the ray
//                               arithmetic of cednerf_generate_rays, op by op).
// No host read anywhere: the threshold, the remaining count and the error flag live in a 32-byte state block.
#include "common.cuh"

namespace {

struct SelectState {
  uint32_t prefix;    // the decided (high) bits of the threshold key
  uint32_t mask;      // which bits are decided
  long long k_rem;    // how many of the keys matching the prefix are still to be taken, counted from the largest
  int err;            // 1: fewer than k positive keys (torch raises "invalid multinomial distribution")
  int pad;
};

__global__ void importance_keys_kernel(const float* __restrict__ w, const int64_t* __restrict__ subset,
                                       const float* __restrict__ noise, int64_t n, uint32_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float wv = w[subset ? subset[i] : i];
  const float q = __fdiv_rn(wv, noise[i]);
  keys[i] = (wv > 0.f && q > 0.f) ? __float_as_uint(q) : 0u;  // zero weight: never drawn (only as an error-flagged filler)
}

__global__ void select_init_kernel(SelectState* st, int64_t k, uint32_t* hist, int nbins) {
  if (threadIdx.x == 0) st->prefix = 0u, st->mask = 0u, st->k_rem = k, st->err = 0, st->pad = 0;
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) hist[b] = 0u;
}

#define SEL_MAX_BINS 2048

__global__ void __launch_bounds__(512) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                         const SelectState* __restrict__ st, int shift, int nbins,
                                                         uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[SEL_MAX_BINS];
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) h[b] = 0u;
  __syncthreads();
  const uint32_t prefix = st->prefix, mask = st->mask, digit = (uint32_t)nbins - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    if ((k & mask) == prefix) atomicAdd(&h[(k >> shift) & digit], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x)
    if (h[b]) atomicAdd(&hist[b], h[b]);
}

// one warp: the digit d with  count(digits > d) < k_rem <= count(digits >= d)  joins the prefix; the histogram is cleared
__global__ void radix_pick_kernel(uint32_t* __restrict__ hist, int nbins, int shift, SelectState* __restrict__ st) {
  const int lane = threadIdx.x, per = nbins / 32;
  const long long k_rem = st->k_rem;
  // lane L owns digits [nbins - (L + 1) per, nbins - L per), i.e. lane 0 the largest
  const int hi = nbins - lane * per;
  long long mine = 0;
  for (int b = hi - per; b < hi; ++b) mine += hist[b];
  long long incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const long long above = incl - mine;  // keys with a digit larger than any of this lane's
  const bool holds = above < k_rem && k_rem <= incl;
  const unsigned who = __ballot_sync(0xffffffffu, holds);
  if (who == 0u) {  // fewer candidates than k: cannot happen after the caller's n >= k check
    if (lane == 0) st->err = 1;
  } else if (lane == __ffs(who) - 1) {
    long long acc = above;
    int d = hi - 1;
    for (; d >= hi - per; --d) {
      if (acc + hist[d] >= k_rem) break;
      acc += hist[d];
    }
    st->prefix |= (uint32_t)d << shift;
    st->mask |= ((uint32_t)nbins - 1u) << shift;
    st->k_rem = k_rem - acc;
    if (shift == 0 && st->prefix == 0u) st->err = 1;  // the k-th largest key is a zero weight
  }
  __syncwarp();
  for (int b = lane; b < nbins; b += 32) hist[b] = 0u;
}

#define SEL_THREADS 256
#define SEL_ITEMS 16

__global__ void __launch_bounds__(SEL_THREADS) select_count_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                   const SelectState* __restrict__ st,
                                                                   int64_t* __restrict__ block_counts) {
  const uint32_t T = st->prefix;
  const int64_t base = ((int64_t)blockIdx.x * SEL_THREADS + threadIdx.x) * SEL_ITEMS;
  int gt = 0, eq = 0;
#pragma unroll
  for (int j = 0; j < SEL_ITEMS; ++j) {
    const int64_t i = base + j;
    if (i < n) {
      const uint32_t k = keys[i];
      gt += k > T, eq += k == T;
    }
  }
  __shared__ int sg[SEL_THREADS / 32], se[SEL_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gt += __shfl_xor_sync(0xffffffffu, gt, o), eq += __shfl_xor_sync(0xffffffffu, eq, o);
  if ((threadIdx.x & 31) == 0) sg[threadIdx.x >> 5] = gt, se[threadIdx.x >> 5] = eq;
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0;
    for (int q = 0; q < SEL_THREADS / 32; ++q) a += sg[q], b += se[q];
    block_counts[2 * blockIdx.x] = a, block_counts[2 * blockIdx.x + 1] = b;
  }
}

// exclusive scan of the (above, equal) pairs of all blocks, in place, by one block (chunks of 1024 with a carry)
__global__ void __launch_bounds__(1024) select_scan_kernel(int64_t* __restrict__ block_counts, int64_t n_blocks) {
  __shared__ int64_t wa[32], wb[32];
  __shared__ int64_t carry_a, carry_b;
  if (threadIdx.x == 0) carry_a = 0, carry_b = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t c0 = 0; c0 < n_blocks; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const int64_t a = i < n_blocks ? block_counts[2 * i] : 0, b = i < n_blocks ? block_counts[2 * i + 1] : 0;
    int64_t ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t va = __shfl_up_sync(0xffffffffu, ia, o), vb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) ia += va, ib += vb;
    }
    if (lane == 31) wa[warp] = ia, wb[warp] = ib;
    __syncthreads();
    if (warp == 0) {
      int64_t ta = wa[lane], tb = wb[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t va = __shfl_up_sync(0xffffffffu, ta, o), vb = __shfl_up_sync(0xffffffffu, tb, o);
        if (lane >= o) ta += va, tb += vb;
      }
      wa[lane] = ta, wb[lane] = tb;  // inclusive over warps
    }
    __syncthreads();
    const int64_t pa = carry_a + (warp ? wa[warp - 1] : 0) + ia - a, pb = carry_b + (warp ? wb[warp - 1] : 0) + ib - b;
    if (i < n_blocks) block_counts[2 * i] = pa, block_counts[2 * i + 1] = pb;
    __syncthreads();
    if (threadIdx.x == 0) carry_a += wa[31], carry_b += wb[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SEL_THREADS) select_scatter_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                     const SelectState* __restrict__ st,
                                                                     const int64_t* __restrict__ block_offsets,
                                                                     const int64_t* __restrict__ subset,
                                                                     int64_t* __restrict__ out) {
  const uint32_t T = st->prefix;
  const long long need = st->k_rem;
  const int64_t base = ((int64_t)blockIdx.x * SEL_THREADS + threadIdx.x) * SEL_ITEMS;
  uint32_t k[SEL_ITEMS];
  int gt = 0, eq = 0;
#pragma unroll
  for (int j = 0; j < SEL_ITEMS; ++j) {
    const int64_t i = base + j;
    k[j] = i < n ? keys[i] : 0u;
    if (i < n) gt += k[j] > T, eq += k[j] == T;
  }
  // exclusive prefix of (gt, eq) over the threads of the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int ig = gt, ie = eq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int vg = __shfl_up_sync(0xffffffffu, ig, o), ve = __shfl_up_sync(0xffffffffu, ie, o);
    if (lane >= o) ig += vg, ie += ve;
  }
  __shared__ int sg[SEL_THREADS / 32], se[SEL_THREADS / 32];
  if (lane == 31) sg[warp] = ig, se[warp] = ie;
  __syncthreads();
  int wg = 0, we = 0;
  for (int q = 0; q < warp; ++q) wg += sg[q], we += se[q];
  long long g_before = block_offsets[2 * blockIdx.x] + wg + ig - gt;
  long long e_before = block_offsets[2 * blockIdx.x + 1] + we + ie - eq;
#pragma unroll
  for (int j = 0; j < SEL_ITEMS; ++j) {
    const int64_t i = base + j;
    if (i >= n) break;
    const bool above = k[j] > T, tie = k[j] == T;
    if (above || (tie && e_before < need)) {
      const long long pos = g_before + (e_before < need ? e_before : need);
      out[pos] = subset ? subset[i] : i;
    }
    g_before += above, e_before += tie;
  }
}

__global__ void importance_batch_kernel(const int64_t* __restrict__ sel, int64_t k, int S, int hsub, int wsub, int W, int H,
                                        const uint8_t* __restrict__ images, const float* __restrict__ c2w, int c2w_rows,
                                        float fx, float fy, float cx, float cy, int opengl,
                                        const float* __restrict__ timestamps, float* __restrict__ origins,
                                        float* __restrict__ viewdirs, float* __restrict__ rgb, float* __restrict__ ts_out,
                                        int64_t* __restrict__ image_id_out, int64_t* __restrict__ pixel_index_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k * S * S) return;
  // x / y / image_id are concatenated over the s x s offsets (ah outer, aw inner), k entries each (:429-437)
  const int sub = (int)(j / k);
  const int64_t i = j - (int64_t)sub * k;
  const int ah = sub / S, aw = sub - ah * S;
  const int64_t index = sel[i], per = (int64_t)hsub * wsub;
  const int64_t image_id = index / per, rem = index - image_id * per;
  const int ysub = (int)(rem / wsub), xsub = (int)(rem - (int64_t)ysub * wsub);
  const int xi = xsub * S + aw, yi = ysub * S + ah;
  const uint8_t* px = images + ((image_id * H + yi) * (int64_t)W + xi) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) rgb[3 * j + c] = __fdiv_rn((float)px[c], 255.0f);
  const float x = (float)xi, y = (float)yi, s = opengl ? -1.f : 1.f;
  const float cd[3] = {__fdiv_rn(__fadd_rn(__fsub_rn(x, cx), 0.5f), fx),
                       __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(y, cy), 0.5f), fy), s), s};
  const float* m = c2w + image_id * (int64_t)(c2w_rows * 4);
  float d[3];
#pragma unroll
  for (int q = 0; q < 3; ++q)
    d[q] = __fadd_rn(__fadd_rn(__fmul_rn(cd[0], m[4 * q]), __fmul_rn(cd[1], m[4 * q + 1])), __fmul_rn(cd[2], m[4 * q + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    origins[3 * j + q] = m[4 * q + 3];
    viewdirs[3 * j + q] = __fdiv_rn(d[q], nrm);
  }
  ts_out[j] = timestamps[image_id];
  if (image_id_out) image_id_out[j] = image_id;
  if (pixel_index_out) pixel_index_out[j] = xi + (int64_t)yi * W + image_id * (int64_t)H * W;
}

}  // namespace

// q = weights[subset[i]] / noise[i] (subset nullable: weights[i]) as order-preserving uint32 keys; noise = Exp(1) draws
CEDNERF_EXPORT int cednerf_importance_keys(const float* weights, const int64_t* subset, const float* exp_noise, int64_t n,
                                           uint32_t* keys, void* stream) {
  CEDNERF_REQUIRE(n >= 0 && weights && exp_noise && keys, "bad arguments");
  if (n == 0) return 0;
  importance_keys_kernel<<<cednerf_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(weights, subset, exp_noise, n, keys);
  return cednerf_check_launch("cednerf_importance_keys");
}

CEDNERF_EXPORT int64_t cednerf_topk_workspace_bytes(int64_t n) {
  const int64_t blocks = (n + SEL_THREADS * SEL_ITEMS - 1) / (SEL_THREADS * SEL_ITEMS);
  return 64 + SEL_MAX_BINS * 4 + (blocks + 1) * 16;
}

// The positions of the k largest keys, ascending (mapped through `subset` when given); ties at the threshold go to the
// lowest positions.  workspace: cednerf_topk_workspace_bytes(n); the int32 at byte 16 is an error flag (1: the
// k-th largest key is zero, i.e. fewer than k positive weights) the caller may read afterwards.
CEDNERF_EXPORT int cednerf_topk_select(const uint32_t* keys, int64_t n, int64_t k, const int64_t* subset, int64_t* out,
                                       void* workspace, void* stream) {
  CEDNERF_REQUIRE(keys && out && workspace && n >= 0 && k >= 0 && k <= n, "bad arguments (k must not exceed n)");
  if (k == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  SelectState* st = (SelectState*)workspace;
  uint32_t* hist = (uint32_t*)((char*)workspace + 64);
  int64_t* block_counts = (int64_t*)((char*)workspace + 64 + SEL_MAX_BINS * 4);
  const int64_t blocks = (n + SEL_THREADS * SEL_ITEMS - 1) / (SEL_THREADS * SEL_ITEMS);
  int launches = 0;
  select_init_kernel<<<1, 256, 0, s>>>(st, k, hist, SEL_MAX_BINS), ++launches;
  const unsigned want = cednerf_blocks(n, 512), cap = (unsigned)cednerf_num_sms() * 4u;
  const unsigned hist_blocks = want < cap ? want : cap;
  const int shifts[3] = {21, 10, 0}, bins[3] = {2048, 2048, 1024};
  for (int p = 0; p < 3; ++p) {
    radix_hist_kernel<<<hist_blocks, 512, 0, s>>>(keys, n, st, shifts[p], bins[p], hist), ++launches;
    radix_pick_kernel<<<1, 32, 0, s>>>(hist, bins[p], shifts[p], st), ++launches;
  }
  select_count_kernel<<<(unsigned)blocks, SEL_THREADS, 0, s>>>(keys, n, st, block_counts), ++launches;
  select_scan_kernel<<<1, 1024, 0, s>>>(block_counts, blocks), ++launches;
  select_scatter_kernel<<<(unsigned)blocks, SEL_THREADS, 0, s>>>(keys, n, st, block_counts, subset, out), ++launches;
  return cednerf_check_launch("cednerf_topk_select", launches);
}

// The batch of fetch_data's training branch from the k drawn cells: ray j = sub * k + i is pixel (xsub s + aw, ysub s + ah)
// of image index / (hsub wsub), sub = ah s + aw.  images uint8 [n_images, H, W, 3]; c2w [n_images, c2w_rows, 4];
// timestamps [n_images].  image_id_out / pixel_index_out nullable.
CEDNERF_EXPORT int cednerf_importance_batch(const int64_t* cells, int64_t k, int subsample, int width, int height,
                                            const uint8_t* images, const float* c2w, int c2w_rows, float fx, float fy,
                                            float cx, float cy, int opengl, const float* timestamps, float* origins,
                                            float* viewdirs, float* rgb, float* timestamps_out, int64_t* image_id_out,
                                            int64_t* pixel_index_out, void* stream) {
  CEDNERF_REQUIRE(cells && k >= 0 && subsample >= 1 && width > 0 && height > 0 && images && c2w && timestamps, "bad arguments");
  CEDNERF_REQUIRE((c2w_rows == 3 || c2w_rows == 4) && origins && viewdirs && rgb && timestamps_out, "bad arguments");
  if (k == 0) return 0;
  const int64_t n = k * subsample * subsample;
  importance_batch_kernel<<<cednerf_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
      cells, k, subsample, height / subsample, width / subsample, width, height, images, c2w, c2w_rows, fx, fy, cx, cy, opengl,
      timestamps, origins, viewdirs, rgb, timestamps_out, image_id_out, pixel_index_out);
  return cednerf_check_launch("cednerf_importance_batch");
}
