// Parameter-free input encodings on the hot path (elementwise, HBM-bound; fused into the MLP operand
// buffers by writing fp16 straight into a strided row of the consumer's input tile):
//   E6 tcnn Frequency(n)            cednerf/model.py:205-213, :316-319, :333-336  
//   E7 tcnn SphericalHarmonics(2)   cednerf/model.py:226-239, used :450-455
//   E4 SinusoidalEncoder            cednerf/encoder.py:28-44
//   E5 SinusoidalEncoderWithExp     cednerf/encoder.py:69-90 (scales :56-61)
#include "common.cuh"

namespace {

// out[s, j] = sin(pi * (2^k x_dim + phase/2)),  dim = j / (2n), k = (j / 2) % n, phase = j % 2
__global__ void frequency_fwd_kernel(const float* __restrict__ x, int n_dims, int64_t n, int n_freq,
                                     __half* __restrict__ out, int out_stride, int pad_to, float pad_value) {
  const int width = n_dims * 2 * n_freq;
  const int cols = pad_to > width ? pad_to : width;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = tid / cols;
  const int j = (int)(tid - s * cols);
  if (s >= n) return;
  float v = pad_value;
  if (j < width) {
    const int dim = j / (2 * n_freq), k = (j >> 1) % n_freq, p = j & 1;
    v = sinpif(x[s * n_dims + dim] * (float)(1 << k) + 0.5f * (float)p);
  }
  out[s * out_stride + j] = __float2half_rn(v);
}

// dx[s, dim] = sum_{k,p} dy[s, j] * 2^k * pi * cos(pi * (2^k x + phase/2))
template <typename GradT>
__global__ void frequency_bwd_kernel(const float* __restrict__ x, int n_dims, int64_t n, int n_freq,
                                     const GradT* __restrict__ dy, int dy_stride, float* __restrict__ dx) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = tid / n_dims;
  const int dim = (int)(tid - s * n_dims);
  if (s >= n) return;
  const float xv = x[s * n_dims + dim];
  float acc = 0.f;
  for (int k = 0; k < n_freq; ++k)
    for (int p = 0; p < 2; ++p) {
      const float sc = (float)(1 << k);
      const float g = (float)dy[s * dy_stride + dim * 2 * n_freq + 2 * k + p];
      acc += g * sc * 3.14159265358979323846f * cospif(xv * sc + 0.5f * (float)p);
    }
  dx[s * n_dims + dim] = acc;
}

// input in [0,1]^3; x,y,z = 2*in-1; [0.2820948, -0.4886025 y, 0.4886025 z, -0.4886025 x]
__global__ void sh2_fwd_kernel(const float* __restrict__ d01, int64_t n, __half* __restrict__ out, int out_stride) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const float x = d01[3 * s] * 2.f - 1.f, y = d01[3 * s + 1] * 2.f - 1.f, z = d01[3 * s + 2] * 2.f - 1.f;
  __half* o = out + s * out_stride;
  o[0] = __float2half_rn(0.28209479177387814f);
  o[1] = __float2half_rn(-0.48860251190291987f * y);
  o[2] = __float2half_rn(0.48860251190291987f * z);
  o[3] = __float2half_rn(-0.48860251190291987f * x);
}

// [t, sin t, sin 2t, sin 4t, sin 8t, sin(t+pi/2), sin(2t+pi/2), ...]           (move_norm == null)
// [t, (sin 2^i t, sin(2^i t + pi/2)) * exp(-(i 2^i) |move|), i = 0..3]          (move_norm != null)
__global__ void time_embed_kernel(const float* __restrict__ t, const float* __restrict__ move_norm, int64_t n,
                                  float* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const float tv = t[s];
  float* o = out + 9 * s;
  o[0] = tv;
  const float half_pi = 1.5707963267948966f;  // rounds to the fp32 value torch's 0.5 * math.pi produces
  if (!move_norm) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float tb = tv * (float)(1 << i);
      o[1 + i] = sinf(tb);
      o[5 + i] = sinf(tb + half_pi);
    }
  } else {
    const float mv = move_norm[s];
    const float scm[4] = {0.f, 2.f, 8.f, 24.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float tb = tv * (float)(1 << i);
      const float att = expf(-1.f * (mv * scm[i]));
      o[1 + 2 * i] = sinf(tb) * att;
      o[2 + 2 * i] = sinf(tb + half_pi) * att;
    }
  }
}

}  // namespace

CEDNERF_EXPORT int cednerf_frequency_fwd(const float* x, int n_dims, int64_t n, int n_frequencies, void* out_f16,
                                         int out_stride, int pad_to, float pad_value, void* stream) {
  CEDNERF_REQUIRE(n >= 0 && n_dims >= 1 && n_frequencies >= 1 && n_frequencies <= 16, "bad sizes");
  const int width = n_dims * 2 * n_frequencies;
  const int cols = pad_to > width ? pad_to : width;
  CEDNERF_REQUIRE(out_stride >= cols, "out_stride too small");
  if (n == 0) return 0;
  frequency_fwd_kernel<<<cednerf_blocks(n * cols, 256), 256, 0, (cudaStream_t)stream>>>(
      x, n_dims, n, n_frequencies, (__half*)out_f16, out_stride, pad_to, pad_value);
  return cednerf_check_launch("cednerf_frequency_fwd");
}

CEDNERF_EXPORT int cednerf_frequency_bwd(const float* x, int n_dims, int64_t n, int n_frequencies, const void* dy,
                                         int dy_stride, int dy_is_f16, float* dx, void* stream) {
  CEDNERF_REQUIRE(n >= 0 && n_dims >= 1 && n_frequencies >= 1 && n_frequencies <= 16, "bad sizes");
  CEDNERF_REQUIRE(dy_stride >= n_dims * 2 * n_frequencies, "dy_stride too small");
  if (n == 0) return 0;
  dim3 grid(cednerf_blocks(n * n_dims, 256));
  if (dy_is_f16)
    frequency_bwd_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n_dims, n, n_frequencies,
                                                                         (const __half*)dy, dy_stride, dx);
  else
    frequency_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, n_dims, n, n_frequencies, (const float*)dy,
                                                                        dy_stride, dx);
  return cednerf_check_launch("cednerf_frequency_bwd");
}

CEDNERF_EXPORT int cednerf_sh2_fwd(const float* d01, int64_t n, void* out_f16, int out_stride, void* stream) {
  CEDNERF_REQUIRE(n >= 0 && out_stride >= 4, "bad sizes");
  if (n == 0) return 0;
  sh2_fwd_kernel<<<cednerf_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(d01, n, (__half*)out_f16, out_stride);
  return cednerf_check_launch("cednerf_sh2_fwd");
}

CEDNERF_EXPORT int cednerf_time_embed(const float* t, const float* move_norm, int64_t n, float* out, void* stream) {
  CEDNERF_REQUIRE(n >= 0, "bad size");
  if (n == 0) return 0;
  time_embed_kernel<<<cednerf_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(t, move_norm, n, out);
  return cednerf_check_launch("cednerf_time_embed");
}
