// Training loss of the reference loop for the canonical flags (train_real.py:369-409): F.mse_loss(rgb, pixels)
// + 1e-3 * acc-entropy (-ae) + 1e-3 * weighted per-sample colour loss (-wr) + latent_losses.mean() (-f), as ONE forward
// and ONE backward launch instead of ~50 element-wise / reduction launches and as many autograd nodes (SURVEY.md §8f N4:
// "losses out" of the hot path).  Traffic: 28 B/ray + 24 B/sample forward, the same again backward - HBM-bound, tiny.
#include "common.cuh"

namespace {

struct LossArgs {
  const float* rgb;      // [R,3] rendered colours
  const float* acc;      // [R]   opacities (nullable: no entropy term)
  const float* pixels;   // [R,3] targets
  const float* rgbs;     // [S,3] per-sample colours (nullable: no per-sample term)
  const float* weights;  // [S]   rendering weights (constants of the loss: the reference detaches them)
  const int64_t* ridx;   // [S]
  const float* latent;   // [R,C] per-ray latent losses (nullable)
  int64_t R, S;          // S: capacity of the per-sample arrays
  const int64_t* S_dev;  // nullable: live sample count on the device
  int C;
  float w_entropy, w_rgbper;
};

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double part[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) part[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (warp == 0) {
    s = lane < (int)(blockDim.x >> 5) ? part[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;  // valid in thread 0
}

// sums[0..3] += (sum (rgb-pix)^2, sum H(1-acc), sum w * |rgbs - pix[ray]|^2, sum latent); grid-stride, 256 threads
__global__ void __launch_bounds__(256) loss_fwd_kernel(LossArgs a, double* __restrict__ sums) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s_mse = 0.0, s_ent = 0.0, s_per = 0.0, s_lat = 0.0;
  for (int64_t r = t0; r < a.R; r += stride) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = a.rgb[3 * r + c] - a.pixels[3 * r + c];
      s_mse += (double)(d * d);
    }
    if (a.acc) {
      const float t = fminf(fmaxf(1.f - a.acc[r], 1e-6f), 1.f - 1e-6f);
      s_ent += (double)(-(t * logf(t) + (1.f - t) * logf(1.f - t)));
    }
  }
  int64_t S = a.S;
  if (a.S_dev) {
    const int64_t v = *a.S_dev;
    S = v < S ? v : S;
  }
  if (a.rgbs)
    for (int64_t s = t0; s < S; s += stride) {
      const int64_t r = a.ridx[s];
      float q = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = a.rgbs[3 * s + c] - a.pixels[3 * r + c];
        q += d * d;
      }
      s_per += (double)(q * a.weights[s]);
    }
  if (a.latent)
    for (int64_t i = t0; i < a.R * a.C; i += stride) s_lat += (double)a.latent[i];
  double v[4] = {s_mse, s_ent, s_per, s_lat};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double s = block_sum(v[k]);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(sums + k, s);
  }
}

__global__ void loss_finish_kernel(LossArgs a, const double* __restrict__ sums, float* __restrict__ loss) {
  const double R = (double)a.R;
  double l = sums[0] / (3.0 * R);
  if (a.acc) l += (double)a.w_entropy * sums[1] / R;
  if (a.rgbs) l += (double)a.w_rgbper * sums[2] / R;
  if (a.latent) l += sums[3] / (R * (double)a.C);
  loss[0] = (float)l;
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(LossArgs a, const float* __restrict__ g_loss, float* __restrict__ d_rgb,
                                                       float* __restrict__ d_acc, float* __restrict__ d_rgbs,
                                                       float* __restrict__ d_latent) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const float g = g_loss[0], inv_r = 1.f / (float)a.R;
  const float k_mse = g * 2.f * inv_r / 3.f, k_ent = g * a.w_entropy * inv_r, k_per = g * a.w_rgbper * 2.f * inv_r;
  for (int64_t r = t0; r < a.R; r += stride) {
    if (d_rgb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) d_rgb[3 * r + c] = k_mse * (a.rgb[3 * r + c] - a.pixels[3 * r + c]);
    }
    if (d_acc) {
      const float t = 1.f - a.acc[r];
      // d/dacc of H(clamp(1 - acc)): log t - log(1 - t) inside the clamp, 0 outside
      d_acc[r] = (t > 1e-6f && t < 1.f - 1e-6f) ? k_ent * (logf(t) - logf(1.f - t)) : 0.f;
    }
  }
  int64_t S = a.S;
  if (a.S_dev) {
    const int64_t v = *a.S_dev;
    S = v < S ? v : S;
  }
  if (d_rgbs)
    for (int64_t s = t0; s < S; s += stride) {
      const int64_t r = a.ridx[s];
      const float w = k_per * a.weights[s];
#pragma unroll
      for (int c = 0; c < 3; ++c) d_rgbs[3 * s + c] = w * (a.rgbs[3 * s + c] - a.pixels[3 * r + c]);
    }
  if (d_latent) {
    const float k_lat = g * inv_r / (float)a.C;
    for (int64_t i = t0; i < a.R * a.C; i += stride) d_latent[i] = k_lat;
  }
}

unsigned loss_grid(const LossArgs& a) {
  const int64_t work = a.R * (a.latent ? a.C : 3) > a.S ? a.R * (a.latent ? a.C : 3) : a.S;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)cednerf_num_sms() * 8;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

// loss[0] = mse(rgb, pixels) + w_entropy * mean H(1 - acc) + w_rgbper * sum_s w_s |rgbs_s - pixels[ray_s]|^2 / R
//           + mean(latent).  acc / rgbs / latent are nullable (term absent).  sums: 4 doubles of workspace.
CEDNERF_EXPORT int cednerf_training_loss_fwd(const float* rgb, const float* acc, const float* pixels, int64_t n_rays,
                                             const float* rgbs, const float* weights, const int64_t* ray_indices,
                                             int64_t n_samples, const float* latent, int n_latent, float w_entropy,
                                             float w_rgbper, double* sums, float* loss, const int64_t* n_device,
                                             void* stream) {
  CEDNERF_REQUIRE(rgb && pixels && n_rays > 0 && sums && loss, "bad arguments");
  CEDNERF_REQUIRE(!rgbs || (weights && ray_indices && n_samples >= 0), "per-sample term needs weights and ray_indices");
  CEDNERF_REQUIRE(!latent || n_latent > 0, "latent term needs its channel count");
  cudaStream_t st = (cudaStream_t)stream;
  LossArgs a{rgb, acc, pixels, n_samples > 0 ? rgbs : nullptr, weights, ray_indices, latent, n_rays,
             rgbs ? n_samples : 0, n_device, n_latent, w_entropy, w_rgbper};
  cudaMemsetAsync(sums, 0, 4 * sizeof(double), st);
  loss_fwd_kernel<<<loss_grid(a), 256, 0, st>>>(a, sums);
  LossArgs fin = a;
  fin.rgbs = rgbs;  // the term exists (as zero) even when this batch has no samples
  loss_finish_kernel<<<1, 1, 0, st>>>(fin, sums, loss);
  return cednerf_check_launch("cednerf_training_loss_fwd", 2);
}

// gradients of the above times g_loss[0] (device scalar: the scaled upstream gradient); any output may be null
CEDNERF_EXPORT int cednerf_training_loss_bwd(const float* g_loss, const float* rgb, const float* acc, const float* pixels,
                                             int64_t n_rays, const float* rgbs, const float* weights,
                                             const int64_t* ray_indices, int64_t n_samples, int n_latent, float w_entropy,
                                             float w_rgbper, float* d_rgb, float* d_acc, float* d_rgbs, float* d_latent,
                                             const int64_t* n_device, void* stream) {
  CEDNERF_REQUIRE(g_loss && rgb && pixels && n_rays > 0, "bad arguments");
  CEDNERF_REQUIRE(!d_acc || acc, "d_acc needs acc");
  CEDNERF_REQUIRE(!d_rgbs || (rgbs && weights && ray_indices), "d_rgbs needs the per-sample inputs");
  LossArgs a{rgb, acc, pixels, rgbs, weights, ray_indices, d_latent ? rgb : nullptr, n_rays, d_rgbs ? n_samples : 0,
             n_device, n_latent > 0 ? n_latent : 1, w_entropy, w_rgbper};
  loss_bwd_kernel<<<loss_grid(a), 256, 0, (cudaStream_t)stream>>>(a, g_loss, d_rgb, d_acc, n_samples > 0 ? d_rgbs : nullptr,
                                                                  d_latent);
  return cednerf_check_launch("cednerf_training_loss_bwd");
}
