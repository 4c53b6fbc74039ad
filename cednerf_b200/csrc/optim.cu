// Fused optimiser step for the training loop of the hot path: what the reference does with
// grad_scaler.step(optimizer) on apex.optimizers.FusedAdam / torch.optim.Adam (train_real.py:252, 267-274, 412-420):
// non-finite check of the scaled gradients, unscale, Adam update of the fp32 master parameters - plus the fp16 working
// copy of the hash table that the next forward reads (the reference re-casts the table on every forward,
// hash_encoder_half.py:381-385) - as one read-only check pass and ONE update pass over up to 8 parameter tensors.
//
// Update pass traffic per parameter: read g, m, v, p (16 B), write m, v, p (12 B) and the fp16 copy (2 B): 30 B,
// HBM-bound (47.9 M table parameters -> 1.44 GB).  torch's path (multi-tensor inf check r/w 8 B, fused Adam 28 B, cast
// 6 B) moves 42 B in three passes.
#include "common.cuh"

#define OPT_MAX_TENSORS 8
#define OPT_CHUNK 4096  // elements per CTA

struct CednerfAdamTensors {
  int n_tensors;
  float* p[OPT_MAX_TENSORS];
  const float* g[OPT_MAX_TENSORS];
  float* m[OPT_MAX_TENSORS];
  float* v[OPT_MAX_TENSORS];
  void* p16[OPT_MAX_TENSORS];  // nullable: fp16 working copy written together with p
  int64_t n[OPT_MAX_TENSORS];
  float lr[OPT_MAX_TENSORS];
  float weight_decay[OPT_MAX_TENSORS];
  int64_t chunk_begin[OPT_MAX_TENSORS + 1];  // filled by the launcher: CTA index range of every tensor
};

namespace {

__device__ __forceinline__ int find_tensor(const CednerfAdamTensors& t, int64_t cta) {
  int k = 0;
  while (k + 1 < t.n_tensors && cta >= t.chunk_begin[k + 1]) ++k;
  return k;
}

// found_inf[0] = 1 when any gradient element is inf / nan (GradScaler's check, without the unscale write-back)
__global__ void __launch_bounds__(256) nonfinite_check_kernel(CednerfAdamTensors t, float* found_inf) {
  const int k = find_tensor(t, blockIdx.x);
  const int64_t base = ((int64_t)blockIdx.x - t.chunk_begin[k]) * OPT_CHUNK;
  const float* g = t.g[k];
  const int64_t n = t.n[k];
  bool bad = false;
  if ((((uintptr_t)g) & 15) == 0 && base + OPT_CHUNK <= n) {
    const float4* g4 = reinterpret_cast<const float4*>(g + base);
#pragma unroll
    for (int j = 0; j < OPT_CHUNK / 4 / 256; ++j) {
      const float4 x = g4[j * 256 + threadIdx.x];
      // x - x is 0 for finite x and nan for inf / nan
      bad = bad || ((x.x - x.x) != 0.f) || ((x.y - x.y) != 0.f) || ((x.z - x.z) != 0.f) || ((x.w - x.w) != 0.f);
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n && i < base + OPT_CHUNK; i += 256) {
      const float x = g[i];
      bad = bad || ((x - x) != 0.f);
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.f;
}

// step[0] += 1 unless a non-finite gradient was found (torch's fused Adam keeps `step` on the device the same way)
__global__ void advance_step_kernel(float* step, const float* found_inf) {
  if (!found_inf || *found_inf == 0.f) *step += 1.f;
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float inv_scale, float lr, float wd,
                                         int adamw, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  g *= inv_scale;
  if (wd != 0.f) {
    if (adamw) p -= lr * wd * p;   // apex FusedAdam default (adam_w_mode)
    else g += wd * p;              // torch.optim.Adam (L2)
  }
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;   // torch._fused_adam_ / _single_tensor_adam arithmetic
  p -= (lr / bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) adam_step_kernel(CednerfAdamTensors t, const float* step, const float* scale_p,
                                                        const float* found_inf, float b1, float b2, float eps, int adamw) {
  const bool skip = found_inf && *found_inf != 0.f;  // GradScaler.step skips the update ...
  const int k = find_tensor(t, blockIdx.x);
  if (skip && !t.p16[k]) return;                     // ... but the fp16 copy is still (re)written so that it is valid
  const int64_t base = ((int64_t)blockIdx.x - t.chunk_begin[k]) * OPT_CHUNK;
  const int64_t n = t.n[k];
  float* p = t.p[k];
  const float* g = t.g[k];
  float* m = t.m[k];
  float* v = t.v[k];
  __half* p16 = reinterpret_cast<__half*>(t.p16[k]);
  const float inv_scale = scale_p ? 1.f / *scale_p : 1.f;  // GradScaler multiplies by the reciprocal of the scale
  const float st = *step;
  const float bc1 = 1.f - powf(b1, st), bc2_sqrt = sqrtf(1.f - powf(b2, st));
  const float lr = t.lr[k], wd = t.weight_decay[k];
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0 &&
                       (!p16 || (((uintptr_t)p16) & 7) == 0);
  if (aligned && base + OPT_CHUNK <= n) {
#pragma unroll
    for (int j = 0; j < OPT_CHUNK / 4 / 256; ++j) {
      const int64_t q = base / 4 + j * 256 + threadIdx.x;
      float4 pp = reinterpret_cast<float4*>(p)[q];
      if (skip) {
        const __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(p16)[q] = o;
        continue;
      }
      const float4 gg = reinterpret_cast<const float4*>(g)[q];
      float4 mm = reinterpret_cast<float4*>(m)[q];
      float4 vv = reinterpret_cast<float4*>(v)[q];
      adam_one(pp.x, gg.x, mm.x, vv.x, inv_scale, lr, wd, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp.y, gg.y, mm.y, vv.y, inv_scale, lr, wd, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp.z, gg.z, mm.z, vv.z, inv_scale, lr, wd, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp.w, gg.w, mm.w, vv.w, inv_scale, lr, wd, adamw, b1, b2, eps, bc1, bc2_sqrt);
      reinterpret_cast<float4*>(p)[q] = pp;
      reinterpret_cast<float4*>(m)[q] = mm;
      reinterpret_cast<float4*>(v)[q] = vv;
      if (p16) {
        const __half2 lo = __floats2half2_rn(pp.x, pp.y), hi = __floats2half2_rn(pp.z, pp.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(p16)[q] = o;
      }
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < n && i < base + OPT_CHUNK; i += 256) {
      float pp = p[i], mm = m[i], vv = v[i];
      if (skip) {
        p16[i] = __float2half_rn(pp);
        continue;
      }
      adam_one(pp, g[i], mm, vv, inv_scale, lr, wd, adamw, b1, b2, eps, bc1, bc2_sqrt);
      p[i] = pp, m[i] = mm, v[i] = vv;
      if (p16) p16[i] = __float2half_rn(pp);
    }
  }
}

int64_t plan_chunks(CednerfAdamTensors& t) {
  int64_t c = 0;
  for (int k = 0; k < t.n_tensors; ++k) {
    t.chunk_begin[k] = c;
    c += (t.n[k] + OPT_CHUNK - 1) / OPT_CHUNK;
  }
  t.chunk_begin[t.n_tensors] = c;
  return c;
}

int check_tensors(const CednerfAdamTensors* t, bool need_state) {
  if (!t || t->n_tensors < 1 || t->n_tensors > OPT_MAX_TENSORS) return 0;
  for (int k = 0; k < t->n_tensors; ++k) {
    if (t->n[k] < 0 || (t->n[k] > 0 && !t->g[k])) return 0;
    if (need_state && t->n[k] > 0 && (!t->p[k] || !t->m[k] || !t->v[k])) return 0;
  }
  return 1;
}

}  // namespace

// found_inf (device float, caller zeroes it) is set to 1 when any gradient element of any tensor is inf / nan.
CEDNERF_EXPORT int cednerf_nonfinite_check(const CednerfAdamTensors* tensors, float* found_inf, void* stream) {
  CEDNERF_REQUIRE(check_tensors(tensors, false) && found_inf, "bad arguments");
  CednerfAdamTensors t = *tensors;
  const int64_t ctas = plan_chunks(t);
  if (ctas == 0) return 0;
  nonfinite_check_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(t, found_inf);
  return cednerf_check_launch("cednerf_nonfinite_check");
}

// One Adam step of every tensor: g is divided by *grad_scale (nullable: 1), the whole step is skipped when
// *found_inf != 0 (nullable), `step` (device float) is advanced first (advance_step != 0) unless skipped; bias
// corrections are formed from it on the device, so nothing is read back by the host.  adam_w_mode: decoupled weight decay (apex default) instead of
// torch.optim.Adam's L2 term; identical when weight_decay == 0 (the reference's setting).
CEDNERF_EXPORT int cednerf_adam_step(const CednerfAdamTensors* tensors, float* step, int advance_step,
                                     const float* grad_scale, const float* found_inf, float beta1, float beta2, float eps,
                                     int adam_w_mode, void* stream) {
  CEDNERF_REQUIRE(check_tensors(tensors, true) && step, "bad arguments");
  CednerfAdamTensors t = *tensors;
  const int64_t ctas = plan_chunks(t);
  if (advance_step) advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, found_inf);
  if (ctas > 0)
    adam_step_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(t, step, grad_scale, found_inf, beta1, beta2, eps,
                                                                       adam_w_mode);
  return cednerf_check_launch("cednerf_adam_step", (ctas > 0 ? 1 : 0) + (advance_step ? 1 : 0));
}
