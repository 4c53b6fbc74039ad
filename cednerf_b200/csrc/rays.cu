// The two steps either side of the hot path: rays in, regulariser out.
//
//  * pixel -> ray generation (datasets/dnerf_3d_video_IS.py:330-358, datasets/dnerf_synthetic.py:196-224, gui.py:43-86):
//    camera_dirs = [(x - cx + 0.5) / fx, (y - cy + 0.5) / fy * s, s] with s = -1 for OpenGL cameras, directions = R
//    camera_dirs, origins = the pose's translation, viewdirs = directions / |directions| - one launch instead of the
//    ~12 element-wise launches (meshgrid, stack, pad, broadcast multiply, sum, norm, divide, reshape) of the reference.
//  * distortion loss (cednerf/losses.py:4-11 -> torch_efficient_distloss.flatten_eff_distloss, Sun et al. 2022, the
//    regulariser of Mip-NeRF 360 in O(N)): per ray, with w the rendering weights, m the interval mid-points and d the
//    interval lengths,  L = sum_i [ d_i w_i^2 / 3 + 2 w_i (m_i W_i - M_i) ],  W_i / M_i = exclusive prefix sums of w and
//    w m along the ray; the loss is sum_rays L / (max ray index + 1).  Backward w.r.t. w only (as the package):
//    dL/dw_i = 2 d_i w_i / 3 + 2 [ m_i (W_i - W'_i) + (M'_i - M_i) ],  W'_i / M'_i = suffix sums.
#include "common.cuh"

namespace {

__global__ void generate_rays_kernel(const int64_t* __restrict__ px, const int64_t* __restrict__ py,
                                     const int64_t* __restrict__ cam, const float* __restrict__ c2w, int c2w_rows,
                                     float fx, float fy, float cx, float cy, int width, int opengl, int64_t n,
                                     float* __restrict__ origins, float* __restrict__ viewdirs, float* __restrict__ directions) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = (float)(px ? px[i] : (i % width)), y = (float)(py ? py[i] : (i / width));
  const float s = opengl ? -1.f : 1.f;
  // element-wise chain of the reference, op by op (no contraction)
  const float c[3] = {__fdiv_rn(__fadd_rn(__fsub_rn(x, cx), 0.5f), fx),
                      __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(y, cy), 0.5f), fy), s), s};
  const float* m = c2w + (cam ? cam[i] : 0) * (int64_t)(c2w_rows * 4);
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    d[k] = __fadd_rn(__fadd_rn(__fmul_rn(c[0], m[4 * k]), __fmul_rn(c[1], m[4 * k + 1])), __fmul_rn(c[2], m[4 * k + 2]));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    origins[3 * i + k] = m[4 * k + 3];
    viewdirs[3 * i + k] = __fdiv_rn(d[k], nrm);
    if (directions) directions[3 * i + k] = d[k];
  }
}

// one warp per ray, lanes walk the ray's samples in chunks of 32 with shuffle scans; a warp keeps an fp64 sum over the rays
// it visits (grid-stride), the block folds its eight warps and writes ONE partial: no atomics (one fp64 atomic per ray onto
// a single address cost 0.27 ms on 2^18 rays), and the sum is the same from run to run
#define DIST_MAX_BLOCKS 2048
__global__ void __launch_bounds__(256) distortion_fwd_kernel(const float* __restrict__ w, const float* __restrict__ t0,
                                                             const float* __restrict__ t1, const int64_t* __restrict__ offsets,
                                                             int64_t n_rays, double* __restrict__ partial_sum,
                                                             unsigned long long* __restrict__ partial_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  unsigned long long last = 0ull;  // 1 + the largest ray index with samples
  for (int64_t ray = (int64_t)blockIdx.x * 8 + warp; ray < n_rays; ray += (int64_t)gridDim.x * 8) {
    const int64_t s0 = offsets[ray], s1 = offsets[ray + 1];
    if (s0 >= s1) continue;
    last = (unsigned long long)(ray + 1);
    float W = 0.f, M = 0.f;  // running exclusive prefixes of w and w m
    for (int64_t base = s0; base < s1; base += 32) {
      const int64_t i = base + lane;
      const bool ok = i < s1;
      const float wi = ok ? w[i] : 0.f, a = ok ? t0[i] : 0.f, b = ok ? t1[i] : 0.f;
      const float mi = (a + b) * 0.5f, di = b - a, wm = wi * mi;
      const float iw = warp_incl_scan_add(wi, lane), iwm = warp_incl_scan_add(wm, lane);
      const float Wp = W + (iw - wi), Mp = M + (iwm - wm);
      if (ok) acc += (double)(di * wi * wi * (1.f / 3.f)) + (double)(2.f * wi * (mi * Wp - Mp));
      W += __shfl_sync(0xffffffffu, iw, 31);
      M += __shfl_sync(0xffffffffu, iwm, 31);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sa[8];
  __shared__ unsigned long long sm[8];
  if (lane == 0) sa[warp] = acc, sm[warp] = last;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    unsigned long long m = 0ull;
    for (int q = 0; q < 8; ++q) {
      t += sa[q];
      m = sm[q] > m ? sm[q] : m;
    }
    partial_sum[blockIdx.x] = t;
    partial_max[blockIdx.x] = m;
  }
}

__global__ void __launch_bounds__(256) distortion_finish_kernel(const double* __restrict__ partial_sum,
                                                                const unsigned long long* __restrict__ partial_max,
                                                                int n_partials, float* __restrict__ loss,
                                                                float* __restrict__ inv_rays) {
  __shared__ double sa[256];
  __shared__ unsigned long long sm[256];
  double t = 0.0;
  unsigned long long m = 0ull;
  for (int q = threadIdx.x; q < n_partials; q += 256) {
    t += partial_sum[q];
    m = partial_max[q] > m ? partial_max[q] : m;
  }
  sa[threadIdx.x] = t, sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sm[threadIdx.x] = sm[threadIdx.x + o] > sm[threadIdx.x] ? sm[threadIdx.x + o] : sm[threadIdx.x];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = (double)sm[0];  // flatten_eff_distloss: n_rays = ray_id.max() + 1
    loss[0] = n > 0 ? (float)(sa[0] / n) : 0.f;
    inv_rays[0] = n > 0 ? (float)(1.0 / n) : 0.f;
  }
}

__global__ void __launch_bounds__(256) distortion_bwd_kernel(const float* __restrict__ w, const float* __restrict__ t0,
                                                             const float* __restrict__ t1, const int64_t* __restrict__ offsets,
                                                             int64_t n_rays, const float* __restrict__ g_loss,
                                                             const float* __restrict__ inv_rays, float* __restrict__ g_w) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const int64_t s0 = offsets[ray], s1 = offsets[ray + 1];
  if (s0 >= s1) return;
  float Wt = 0.f, Mt = 0.f;  // the ray's totals
  for (int64_t base = s0; base < s1; base += 32) {
    const int64_t i = base + lane;
    const bool ok = i < s1;
    const float wi = ok ? w[i] : 0.f;
    const float wm = ok ? wi * ((t0[i] + t1[i]) * 0.5f) : 0.f;
    Wt += warp_sum(wi);
    Mt += warp_sum(wm);
  }
  const float scale = g_loss[0] * inv_rays[0];
  float W = 0.f, M = 0.f;
  for (int64_t base = s0; base < s1; base += 32) {
    const int64_t i = base + lane;
    const bool ok = i < s1;
    const float wi = ok ? w[i] : 0.f, a = ok ? t0[i] : 0.f, b = ok ? t1[i] : 0.f;
    const float mi = (a + b) * 0.5f, di = b - a, wm = wi * mi;
    const float iw = warp_incl_scan_add(wi, lane), iwm = warp_incl_scan_add(wm, lane);
    const float Wp = W + (iw - wi), Mp = M + (iwm - wm);
    const float Ws = Wt - (Wp + wi), Ms = Mt - (Mp + wm);
    if (ok) g_w[i] = scale * (di * 2.f * wi * (1.f / 3.f) + 2.f * (mi * (Wp - Ws) + (Ms - Mp)));
    W += __shfl_sync(0xffffffffu, iw, 31);
    M += __shfl_sync(0xffffffffu, iwm, 31);
  }
}

}  // namespace

// pixel -> ray (see the file header).  px / py (nullable): pixel coordinates per ray, else ray i is pixel (i % width,
// i / width) of one frame; cam (nullable): pose index per ray into c2w [n_cams, c2w_rows (3 or 4), 4], else pose 0.
// directions (nullable): the un-normalised directions.
CEDNERF_EXPORT int cednerf_generate_rays(const int64_t* px, const int64_t* py, const int64_t* cam, const float* c2w,
                                         int c2w_rows, float fx, float fy, float cx, float cy, int width, int opengl,
                                         int64_t n, float* origins, float* viewdirs, float* directions, void* stream) {
  CEDNERF_REQUIRE(n >= 0 && c2w && (c2w_rows == 3 || c2w_rows == 4) && width > 0 && origins && viewdirs, "bad arguments");
  CEDNERF_REQUIRE((px == nullptr) == (py == nullptr), "px and py go together");
  if (n == 0) return 0;
  generate_rays_kernel<<<cednerf_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(px, py, cam, c2w, c2w_rows, fx, fy, cx, cy,
                                                                                  width, opengl, n, origins, viewdirs, directions);
  return cednerf_check_launch("cednerf_generate_rays");
}

// distortion loss of packed samples (offsets [n_rays + 1]); work: cednerf_distortion_workspace_bytes() (block partials),
// loss / inv_rays: 1 float each (inv_rays feeds the backward).  Empty rays contribute nothing; n_rays of the normalisation
// = last ray with samples + 1.
CEDNERF_EXPORT int64_t cednerf_distortion_workspace_bytes(void) { return (int64_t)DIST_MAX_BLOCKS * 16; }

CEDNERF_EXPORT int cednerf_distortion_fwd(const float* weights, const float* t_starts, const float* t_ends,
                                          const int64_t* offsets, int64_t n_rays, void* work, float* loss, float* inv_rays,
                                          void* stream) {
  CEDNERF_REQUIRE(n_rays >= 0 && offsets && work && loss && inv_rays, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  double* psum = (double*)work;
  unsigned long long* pmax = (unsigned long long*)work + DIST_MAX_BLOCKS;
  int launches = 1, blocks = 0;
  if (n_rays > 0) {
    const int64_t want = (n_rays + 7) / 8, cap = (int64_t)cednerf_num_sms() * 8;
    blocks = (int)(want < cap ? want : cap);
    if (blocks > DIST_MAX_BLOCKS) blocks = DIST_MAX_BLOCKS;
    distortion_fwd_kernel<<<blocks, 256, 0, st>>>(weights, t_starts, t_ends, offsets, n_rays, psum, pmax);
    ++launches;
  }
  distortion_finish_kernel<<<1, 256, 0, st>>>(psum, pmax, blocks, loss, inv_rays);
  return cednerf_check_launch("cednerf_distortion_fwd", launches);
}

// g_weights[i] = g_loss[0] * dL/dw_i for every sample of a non-empty ray (the caller zero-fills the rest if it matters)
CEDNERF_EXPORT int cednerf_distortion_bwd(const float* weights, const float* t_starts, const float* t_ends,
                                          const int64_t* offsets, int64_t n_rays, const float* g_loss, const float* inv_rays,
                                          float* g_weights, void* stream) {
  CEDNERF_REQUIRE(n_rays >= 0 && offsets && g_loss && inv_rays && g_weights, "bad arguments");
  if (n_rays == 0) return 0;
  distortion_bwd_kernel<<<cednerf_blocks(n_rays * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      weights, t_starts, t_ends, offsets, n_rays, g_loss, inv_rays, g_weights);
  return cednerf_check_launch("cednerf_distortion_bwd");
}
