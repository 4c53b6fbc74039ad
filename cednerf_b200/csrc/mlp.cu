// K3 — fully fused 64-wide MLP, forward and backward, on tcgen05 tensor cores (sm_100a).
//
// Same network family as the tcnn.Network / NetworkWithInputEncoding modules the reference builds at
// cednerf/model.py:200-222 (xyz_wrap), :280-290 (mlp_base), :292-309 (mlp_head), :312-344 (predictors):
// FullyFusedMLP, 64 neurons, ReLU hidden, no output activation, no bias, fp16 weights/activations
// (SURVEY.md Appendix B).  Numerics shared with oracle/tcnn_ref.py: fp32 accumulate, hidden activations and
// inter-layer gradients rounded to fp16, weight gradients fp32.
//
// Design (one CTA = 128 threads = one 128-sample tile at a time, persistent over tiles):
//   * every operand tile lives in shared memory as [rows][128 B] with the 128-byte swizzle, so the SAME
//     bytes serve as a K-major operand (rows = M or N, 64 fp16 along K) and as an MN-major operand
//     (rows = K, 64 fp16 along M or N).  Forward uses W as K-major B; dgrad re-reads the same W image as
//     MN-major B (no transposed copy); wgrad reads the gradient tile and the activation tile MN-major
//     with the 128 samples as K.
//   * accumulators live in TMEM: one 128x64 fp32 tile for the layer chain, plus (backward) one 64xK tile
//     per layer that accumulates dW across all tiles a CTA processes and is flushed once with fp32 atomics.
//   * one thread issues tcgen05.mma; completion comes back through tcgen05.commit on an mbarrier; the four
//     warps then pull their TMEM lane quarter with tcgen05.ld, apply ReLU / the ReLU mask, round to fp16
//     and write the next operand tile straight back to shared memory - activations never touch HBM
//     between layers (the hidden tiles are streamed out only when the backward pass needs them).
#include "tc05.cuh"

namespace {

struct MlpFwdArgs {
  const __half* x;       // [n, dim_in[0]]
  const uint8_t* image;  // swizzled fp16 weight images
  __half* out;           // [n, dim_out[last]]
  __half* hidden;        // [n_layers-1][n][64] post-ReLU activations, or null
  int64_t n;
  CednerfMlpDesc d;
};

__global__ void __launch_bounds__(MLP_TILE) mlp_fwd_kernel(MlpFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                                            // weight images (<= 5 x 8 KB)
  uint8_t* abuf[2] = {smem + MLP_MAX_LAYERS * 8192, smem + MLP_MAX_LAYERS * 8192 + MLP_TILE_BYTES};
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const CednerfMlpDesc& d = a.d;
  for (int q = tid; q < d.image_bytes / 16; q += blockDim.x)
    reinterpret_cast<uint4*>(wsm)[q] = __ldg(reinterpret_cast<const uint4*>(a.image) + q);
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_warp = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  const int L = d.n_layers;
  const int64_t n_tiles = (a.n + MLP_TILE - 1) / MLP_TILE;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * MLP_TILE;
    const int rows_valid = (int)((a.n - row0) < MLP_TILE ? (a.n - row0) : MLP_TILE);
    load_tile(abuf[0], a.x + row0 * d.dim_in[0], d.dim_in[0], rows_valid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    int cur = 0;
    for (int l = 0; l < L; ++l) {
      const int K = d.dim_in[l], N = d.dim_out[l];
      if (tid == 0) {
        tc_fence_after();
        const uint64_t ad = make_desc(smem_u32(abuf[cur]), 1, 64);
        const uint64_t bd = make_desc(smem_u32(wsm + d.image_off[l]), 1, 64);
        const uint32_t id = make_idesc(128, N, 0, 0);
        for (int k = 0; k < K / 16; ++k) umma(tmem_base, ad + 2 * k, bd + 2 * k, id, k > 0);
        umma_commit(&bar);
      }
      // stream the previous hidden tile out while the tensor core works on this layer
      if (l > 0 && a.hidden)
        store_tile(abuf[cur], a.hidden + ((int64_t)(l - 1) * a.n + row0) * 64, 64, rows_valid);
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (l < L - 1) {
        uint8_t* nxt = abuf[cur ^ 1];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t r[16];
          tmem_ld16(tmem_warp + cb * 16, r);
          tmem_ld_wait();
          uint32_t p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = pack_relu_h2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          *reinterpret_cast<uint4*>(nxt + swz(tid, 2 * cb)) = make_uint4(p[0], p[1], p[2], p[3]);
          *reinterpret_cast<uint4*>(nxt + swz(tid, 2 * cb + 1)) = make_uint4(p[4], p[5], p[6], p[7]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        cur ^= 1;
      } else {
        __half* orow = a.out + (row0 + tid) * N;
        for (int cb = 0; cb < N / 16; ++cb) {
          uint32_t r[16];
          tmem_ld16(tmem_warp + cb * 16, r);
          tmem_ld_wait();
          if (tid < rows_valid) {
            uint32_t p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) p[j] = pack_h2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            reinterpret_cast<uint4*>(orow)[2 * cb] = make_uint4(p[0], p[1], p[2], p[3]);
            reinterpret_cast<uint4*>(orow)[2 * cb + 1] = make_uint4(p[4], p[5], p[6], p[7]);
          }
        }
        tc_fence_before();
        __syncthreads();  // TMEM and both activation tiles are free for the next tile
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

struct MlpBwdArgs {
  const __half* x;        // [n, dim_in[0]] saved input
  const __half* hidden;   // [n_layers-1][n][64] saved post-ReLU activations
  const __half* d_out;    // [n, dim_out[last]]
  const uint8_t* image;
  void* d_x;              // [n, dim_in[0]] fp32 or fp16, or null
  int dx_is_f32;
  float* d_params;        // flat fp32, same layout as params (atomically accumulated), or null
  __half* d_hidden;       // [n_layers-1][n][64] gradients w.r.t. the hidden activations after the ReLU mask, or null
  int64_t n;
  uint32_t tmem_cols;
  CednerfMlpDesc d;
};

__global__ void __launch_bounds__(MLP_TILE) mlp_bwd_kernel(MlpBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;
  uint8_t* gbuf[2] = {smem + MLP_MAX_LAYERS * 8192, smem + MLP_MAX_LAYERS * 8192 + MLP_TILE_BYTES};
  uint8_t* ibuf[2] = {smem + MLP_MAX_LAYERS * 8192 + 2 * MLP_TILE_BYTES, smem + MLP_MAX_LAYERS * 8192 + 3 * MLP_TILE_BYTES};
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const CednerfMlpDesc& d = a.d;
  for (int q = tid; q < d.image_bytes / 16; q += blockDim.x)
    reinterpret_cast<uint4*>(wsm)[q] = __ldg(reinterpret_cast<const uint4*>(a.image) + q);
  if (warp == 0) tmem_alloc(&tmem_base_s, a.tmem_cols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_warp = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint32_t phase = 0;
  const int L = d.n_layers;
  const int64_t n_tiles = (a.n + MLP_TILE - 1) / MLP_TILE;
  const bool want_dw = a.d_params != nullptr;
  int64_t iter = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
    const int64_t row0 = tile * MLP_TILE;
    const int rows_valid = (int)((a.n - row0) < MLP_TILE ? (a.n - row0) : MLP_TILE);
    load_tile_async(gbuf[0], a.d_out + row0 * d.dim_out[L - 1], d.dim_out[L - 1], rows_valid);
    if (L > 1) load_tile_async(ibuf[0], a.hidden + ((int64_t)(L - 2) * a.n + row0) * 64, 64, rows_valid);
    else load_tile_async(ibuf[0], a.x + row0 * d.dim_in[0], d.dim_in[0], rows_valid);
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    int cur = 0;
    for (int l = L - 1; l >= 0; --l) {
      const int K_in = d.dim_in[l], N_out = d.dim_out[l];
      const bool need_dgrad = (l > 0) || (a.d_x != nullptr);
      if (tid == 0) {
        tc_fence_after();
        if (want_dw) {
          // dW_l[out, in] += G^T[out, samples] * I[samples, in]: both operands MN-major, K = 128 samples
          const uint64_t ad = make_desc(smem_u32(gbuf[cur]), 64, 64);
          const uint64_t bd = make_desc(smem_u32(ibuf[cur]), 64, 64);
          const uint32_t id = make_idesc(64, K_in, 1, 1);
          const uint32_t acc = tmem_base + 64u * (uint32_t)(l + 1);
          for (int k = 0; k < MLP_TILE / 16; ++k) umma(acc, ad + 128 * k, bd + 128 * k, id, (iter > 0) || (k > 0));
        }
        if (need_dgrad) {
          // dI[samples, in] = G[samples, out] * W_l[out, in]: A K-major, B = the forward weight image read MN-major
          const uint64_t ad = make_desc(smem_u32(gbuf[cur]), 1, 64);
          const uint64_t bd = make_desc(smem_u32(wsm + d.image_off[l]), 64, 64);
          const uint32_t id = make_idesc(128, K_in, 0, 1);
          for (int k = 0; k < N_out / 16; ++k) umma(tmem_base, ad + 2 * k, bd + 128 * k, id, k > 0);
        }
        umma_commit(&bar);
      }
      // prefetch the next (earlier) layer's input tile while the tensor core runs
      if (l > 0) {
        if (l > 1) load_tile_async(ibuf[cur ^ 1], a.hidden + ((int64_t)(l - 2) * a.n + row0) * 64, 64, rows_valid);
        else load_tile_async(ibuf[cur ^ 1], a.x + row0 * d.dim_in[0], d.dim_in[0], rows_valid);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (l > 0) {
        // G_{l-1} = dI (*) [H_{l-1} > 0], rounded to fp16, written as the next gradient tile
        uint8_t* nxt = gbuf[cur ^ 1];
        const uint8_t* act = ibuf[cur];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t r[16];
          tmem_ld16(tmem_warp + cb * 16, r);
          tmem_ld_wait();
          const uint4 m0 = *reinterpret_cast<const uint4*>(act + swz(tid, 2 * cb));
          const uint4 m1 = *reinterpret_cast<const uint4*>(act + swz(tid, 2 * cb + 1));
          const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          uint32_t p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // ReLU mask from the saved post-ReLU activations: one packed compare (0xffff per half that is > 0) and
            // one AND on the packed gradient - exact zeroing, three instructions per pair of elements
            const unsigned keep = __hgt2_mask(*reinterpret_cast<const __half2*>(&mw[j]), __float2half2_rn(0.f));
            p[j] = pack_h2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])) & keep;
          }
          *reinterpret_cast<uint4*>(nxt + swz(tid, 2 * cb)) = make_uint4(p[0], p[1], p[2], p[3]);
          *reinterpret_cast<uint4*>(nxt + swz(tid, 2 * cb + 1)) = make_uint4(p[4], p[5], p[6], p[7]);
          if (a.d_hidden && tid < rows_valid) {
            uint4* dst = reinterpret_cast<uint4*>(a.d_hidden + ((int64_t)(l - 1) * a.n + row0 + tid) * 64 + cb * 16);
            dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
            dst[1] = make_uint4(p[4], p[5], p[6], p[7]);
          }
        }
      } else if (a.d_x) {
        for (int cb = 0; cb < K_in / 16; ++cb) {
          uint32_t r[16];
          tmem_ld16(tmem_warp + cb * 16, r);
          tmem_ld_wait();
          if (tid < rows_valid) {
            if (a.dx_is_f32) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.d_x) + (row0 + tid) * K_in + cb * 16);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            } else {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.d_x) + (row0 + tid) * K_in + cb * 16);
              uint32_t p[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) p[j] = pack_h2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
              dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
              dst[1] = make_uint4(p[4], p[5], p[6], p[7]);
            }
          }
        }
      }
      cp_async_wait_all();  // the prefetched next-layer tile has landed
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      cur ^= 1;
    }
  }

  // flush the per-CTA weight-gradient accumulators (M = 64: row m lives in lane (m % 16) of warp m / 16)
  if (want_dw && iter > 0) {
    tc_fence_after();
    for (int l = 0; l < L; ++l) {
      const int K_in = d.dim_in[l], N_out = d.dim_out[l];
      const int m = warp * 16 + lane;
      for (int cb = 0; cb < K_in / 16; ++cb) {
        uint32_t r[16];
        tmem_ld16(tmem_warp + 64u * (uint32_t)(l + 1) + cb * 16, r);
        tmem_ld_wait();
        if (lane < 16 && m < N_out) {
          float* dst = a.d_params + d.param_off[l] + m * K_in + cb * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// flat fp32 params ([dim_out, dim_in] row-major per layer) -> swizzled fp16 images
__global__ void mlp_pack_kernel(const float* __restrict__ params, CednerfMlpDesc d, uint8_t* __restrict__ image) {
  const int total_chunks = d.image_bytes / 16;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total_chunks; q += gridDim.x * blockDim.x) {
    const int byte = q * 16;
    int l = 0;
    while (l + 1 < d.n_layers && byte >= d.image_off[l + 1]) ++l;
    const int local = byte - d.image_off[l];
    const int r = local / MLP_ROW_BYTES;                    // output neuron
    const int c = ((local % MLP_ROW_BYTES) >> 4) ^ (r & 7);  // logical 16-byte chunk held at this position
    uint32_t p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k0 = c * 8 + 2 * j;
      const float v0 = k0 < d.dim_in[l] ? params[d.param_off[l] + r * d.dim_in[l] + k0] : 0.f;
      const float v1 = k0 + 1 < d.dim_in[l] ? params[d.param_off[l] + r * d.dim_in[l] + k0 + 1] : 0.f;
      p[j] = pack_h2(v0, v1);
    }
    reinterpret_cast<uint4*>(image)[q] = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

int check_desc(const CednerfMlpDesc* d) {
  if (!d || d->n_layers < 1 || d->n_layers > MLP_MAX_LAYERS) return 0;
  int off = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    const int in = d->dim_in[l], out = d->dim_out[l];
    if (in < 16 || in > 64 || (in % 16) || out < 16 || out > 64 || (out % 16)) return 0;
    if (l > 0 && in != 64) return 0;
    if (l < d->n_layers - 1 && out != 64) return 0;
    if (d->image_off[l] != off) return 0;
    off += out * MLP_ROW_BYTES;
  }
  return d->image_bytes == off;
}

constexpr int FWD_SMEM = MLP_MAX_LAYERS * 8192 + 2 * MLP_TILE_BYTES + 1024;
constexpr int BWD_SMEM = MLP_MAX_LAYERS * 8192 + 4 * MLP_TILE_BYTES + 1024;

}  // namespace

CEDNERF_EXPORT int cednerf_mlp_pack_weights(const float* params, const CednerfMlpDesc* desc, void* image, void* stream) {
  CEDNERF_REQUIRE(check_desc(desc), "bad MLP descriptor");
  mlp_pack_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(params, *desc, (uint8_t*)image);
  return cednerf_check_launch("cednerf_mlp_pack_weights");
}

CEDNERF_EXPORT int cednerf_mlp_fwd(const void* x_f16, const void* weight_image, const CednerfMlpDesc* desc, int64_t n,
                                   void* out_f16, void* hidden_f16, void* stream) {
  CEDNERF_REQUIRE(check_desc(desc), "bad MLP descriptor");
  CEDNERF_REQUIRE(n >= 0, "bad size");
  if (n == 0) return 0;
  static CednerfOncePerDevice configured;
  if (int e = cednerf_opt_in_smem(mlp_fwd_kernel, FWD_SMEM, configured, "cednerf_mlp_fwd")) return e;
  MlpFwdArgs a{(const __half*)x_f16, (const uint8_t*)weight_image, (__half*)out_f16, (__half*)hidden_f16, n, *desc};
  const int64_t tiles = (n + MLP_TILE - 1) / MLP_TILE;
  const int64_t max_ctas = (int64_t)cednerf_num_sms() * 3;
  mlp_fwd_kernel<<<(unsigned)(tiles < max_ctas ? tiles : max_ctas), MLP_TILE, FWD_SMEM, (cudaStream_t)stream>>>(a);
  return cednerf_check_launch("cednerf_mlp_fwd");
}

CEDNERF_EXPORT int cednerf_mlp_bwd(const void* x_f16, const void* hidden_f16, const void* d_out_f16,
                                   const void* weight_image, const CednerfMlpDesc* desc, int64_t n, void* d_x,
                                   int dx_is_f32, float* d_params, void* d_hidden_f16, void* stream) {
  CEDNERF_REQUIRE(check_desc(desc), "bad MLP descriptor");
  CEDNERF_REQUIRE(n >= 0 && (desc->n_layers == 1 || hidden_f16), "bad arguments");
  if (n == 0) return 0;
  static CednerfOncePerDevice configured;
  if (int e = cednerf_opt_in_smem(mlp_bwd_kernel, BWD_SMEM, configured, "cednerf_mlp_bwd")) return e;
  uint32_t cols = 64u * (uint32_t)(desc->n_layers + 1), alloc = 64;
  while (alloc < cols) alloc <<= 1;
  MlpBwdArgs a{(const __half*)x_f16, (const __half*)hidden_f16, (const __half*)d_out_f16, (const uint8_t*)weight_image,
               d_x, dx_is_f32, d_params, (__half*)d_hidden_f16, n, alloc, *desc};
  const int64_t tiles = (n + MLP_TILE - 1) / MLP_TILE;
  const int64_t max_ctas = (int64_t)cednerf_num_sms() * (alloc <= 256 ? 2 : 1);
  mlp_bwd_kernel<<<(unsigned)(tiles < max_ctas ? tiles : max_ctas), MLP_TILE, BWD_SMEM, (cudaStream_t)stream>>>(a);
  return cednerf_check_launch("cednerf_mlp_bwd");
}
