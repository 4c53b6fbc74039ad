// Data-parallel optimiser step over NVLink peer memory (the reference is single-GPU,
// train_real.py:81 hard-codes cuda:0, so there is no reference collective to mirror).
//
// One process per GPU.  Every rank keeps its hash-table gradient, the fp32 master table and the fp16 working copy in
// buffers that its peers map through CUDA IPC.  After the backward pass
//
//   barrier  ->  fused kernel  ->  barrier
//
// replaces "all-reduce the 191 MB gradient, then Adam on the whole table on every rank":  rank r owns elements
// [lo_r, hi_r) of the table.  Its fused kernel reads that slice of EVERY rank's gradient buffer straight over NVLink
// (reduce-scatter: (N-1)/N x 4 B per owned element inbound), sums in rank order (so the result does not depend on who
// computes it), applies the unscale + Adam update to its slice of the master parameters - the Adam moments exist only
// for the owned slice - and stores the updated fp32 value and its fp16 copy into all N replicas (all-gather, 6 B per
// element and peer outbound).  No gradient, parameter or moment makes a second trip through HBM: per owned element the
// kernel moves 4 N B of gradients in, 16 B of local state in/out and 6 N B of parameters out, against
// (2 (N-1)/N x 4 B on the wire + 30 B of Adam traffic) x N elements for all-reduce + replicated Adam.
//
// The barriers are flag exchanges in the same peer memory (one 32-bit epoch per rank pair, st.release.sys /
// ld.acquire.sys), enqueued on the training stream like any other kernel: no host synchronisation, no NCCL call.
#include <string.h>

#include "common.cuh"

#define DP_MAX_RANKS 8
#define DP_CHUNK 4096  // elements per CTA of the fused kernel

// Control block, one per rank, in peer-visible memory.
struct CednerfDpCtrl {
  uint32_t arrive[DP_MAX_RANKS];  // arrive[q] = last barrier epoch rank q has signalled to this rank
  float found_inf;                // this rank's local "a gradient is inf / nan" flag (GradScaler)
  uint32_t timed_out;             // set when a barrier gave up waiting (a peer died): surfaced by the host
};

struct CednerfDpPeers {
  int world, rank;
  CednerfDpCtrl* ctrl[DP_MAX_RANKS];  // every rank's control block as mapped in THIS process (ctrl[rank] is local)
};

struct CednerfDpAdam {
  int world, rank;
  const float* grad[DP_MAX_RANKS];  // every rank's gradient buffer (same layout on all ranks), peer-mapped
  int n_out;                        // number of replicas to update: world (broadcast) or 1 (local only)
  float* p32_out[DP_MAX_RANKS];     // fp32 parameter replicas (entry 0 must be the local one)
  void* p16_out[DP_MAX_RANKS];      // fp16 working copies, nullable
  float* m;                         // Adam moments of the OWNED range, indexed from lo
  float* v;
  int64_t lo, hi;                   // owned element range; lo is a multiple of 4
  float lr, weight_decay, grad_div; // the summed gradient is divided by grad_div (world for an average)
};

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Thread q tells rank q "rank `rank` has reached epoch" and waits until rank q has told us the same.  Everything this
// rank enqueued before the barrier (its gradient kernels, its parameter stores into peer replicas) has completed by
// stream order and is fenced at system scope; everything enqueued after it may read what the peers wrote before theirs.
__global__ void dp_barrier_kernel(CednerfDpPeers p, uint32_t epoch, uint64_t timeout_ns) {
  const int q = threadIdx.x;
  if (q >= p.world) return;
  __threadfence_system();
  if (q != p.rank) st_release_sys(&p.ctrl[q]->arrive[p.rank], epoch);
  if (q == p.rank) return;
  const uint32_t* mine = &p.ctrl[p.rank]->arrive[q];
  const uint64_t t0 = global_ns();
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    __nanosleep(200);
    if (global_ns() - t0 > timeout_ns) {  // a peer is gone: report, do not hang the device
      p.ctrl[p.rank]->timed_out = epoch;
      break;
    }
  }
  __threadfence_system();
}

// found_out = OR over ranks of their local found-inf flags; the shared step counter advances unless the step is skipped.
__global__ void dp_found_kernel(CednerfDpPeers p, float* found_out, float* step) {
  float f = 0.f;
  for (int r = 0; r < p.world; ++r)
    if (*reinterpret_cast<const volatile float*>(&p.ctrl[r]->found_inf) != 0.f) f = 1.f;
  if (found_out) *found_out = f;
  if (step && f == 0.f) *step += 1.f;
}

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// identical arithmetic to optim.cu::adam_one (torch._fused_adam_ / _single_tensor_adam)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float g_mul, float lr, float wd, int adamw,
                                         float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  g *= g_mul;
  if (wd != 0.f) {
    if (adamw) p -= lr * wd * p;
    else g += wd * p;
  }
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p -= (lr / bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) dp_adam_kernel(CednerfDpAdam a, const float* step, const float* scale_p,
                                                      const float* found_inf, float b1, float b2, float eps, int adamw) {
  if (found_inf && *found_inf != 0.f) return;  // GradScaler skips the step on every rank alike; replicas stay as they are
  const int64_t base = a.lo + (int64_t)blockIdx.x * DP_CHUNK;
  // GradScaler multiplies by the reciprocal of the scale; the average over ranks is a second exact-or-rounded factor
  const float inv_scale = scale_p ? 1.f / *scale_p : 1.f;
  const float g_mul = a.grad_div != 1.f ? inv_scale / a.grad_div : inv_scale;
  const float st = *step;
  const float bc1 = 1.f - powf(b1, st), bc2_sqrt = sqrtf(1.f - powf(b2, st));
  const float* p_in = a.p32_out[0];
  if (base + DP_CHUNK <= a.hi) {
#pragma unroll
    for (int j = 0; j < DP_CHUNK / 4 / 256; ++j) {
      const int64_t e = base + (int64_t)(j * 256 + threadIdx.x) * 4;
      float4 g = ld_stream4(a.grad[0] + e);
#pragma unroll
      for (int r = 1; r < DP_MAX_RANKS; ++r)
        if (r < a.world) {
          const float4 x = ld_stream4(a.grad[r] + e);
          g.x += x.x, g.y += x.y, g.z += x.z, g.w += x.w;
        }
      float4 pp = *reinterpret_cast<const float4*>(p_in + e);
      float4 mm = *reinterpret_cast<const float4*>(a.m + (e - a.lo));
      float4 vv = *reinterpret_cast<const float4*>(a.v + (e - a.lo));
      adam_one(pp.x, g.x, mm.x, vv.x, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp.y, g.y, mm.y, vv.y, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp.z, g.z, mm.z, vv.z, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp.w, g.w, mm.w, vv.w, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      *reinterpret_cast<float4*>(a.m + (e - a.lo)) = mm;
      *reinterpret_cast<float4*>(a.v + (e - a.lo)) = vv;
      const __half2 h0 = __floats2half2_rn(pp.x, pp.y), h1 = __floats2half2_rn(pp.z, pp.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&h0);
      o.y = *reinterpret_cast<const uint32_t*>(&h1);
#pragma unroll
      for (int r = 0; r < DP_MAX_RANKS; ++r)
        if (r < a.n_out) {
          *reinterpret_cast<float4*>(a.p32_out[r] + e) = pp;
          if (a.p16_out[r]) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.p16_out[r]) + e) = o;
        }
    }
  } else {
    for (int64_t e = base + threadIdx.x; e < a.hi && e < base + DP_CHUNK; e += 256) {
      float g = a.grad[0][e];
      for (int r = 1; r < a.world; ++r) g += a.grad[r][e];
      float pp = p_in[e], mm = a.m[e - a.lo], vv = a.v[e - a.lo];
      adam_one(pp, g, mm, vv, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      a.m[e - a.lo] = mm, a.v[e - a.lo] = vv;
      for (int r = 0; r < a.n_out; ++r) {
        a.p32_out[r][e] = pp;
        if (a.p16_out[r]) reinterpret_cast<__half*>(a.p16_out[r])[e] = __float2half_rn(pp);
      }
    }
  }
  __threadfence_system();  // the replica stores are performed before this rank's next barrier signal
}

}  // namespace

// ---- peer-visible memory (CUDA IPC): the caller exchanges the 64-byte handles between the ranks -------------------
CEDNERF_EXPORT int cednerf_peer_alloc(int64_t bytes, void** ptr) {
  CEDNERF_REQUIRE(bytes > 0 && ptr, "bad arguments");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) {
    cednerf_set_error("cednerf_peer_alloc: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

CEDNERF_EXPORT int cednerf_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    cednerf_set_error("cednerf_peer_free: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

CEDNERF_EXPORT int cednerf_ipc_export(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CEDNERF_REQUIRE(ptr && handle64, "bad arguments");
  cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
  if (e != cudaSuccess) {
    cednerf_set_error("cednerf_ipc_export: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// Maps a peer's allocation into this process on the CURRENT device (peer access between the two devices is enabled
// lazily by the driver); *ptr is the mapped base address.
CEDNERF_EXPORT int cednerf_ipc_open(const void* handle64, void** ptr) {
  CEDNERF_REQUIRE(handle64 && ptr, "bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cednerf_set_error("cednerf_ipc_open: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

CEDNERF_EXPORT int cednerf_ipc_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    cednerf_set_error("cednerf_ipc_close: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

CEDNERF_EXPORT int64_t cednerf_dp_ctrl_bytes(void) { return (int64_t)sizeof(CednerfDpCtrl); }

// ---- stream-ordered cross-rank barrier -----------------------------------------------------------------------------
// `epoch` must increase by one per barrier, identically on every rank.  timeout_ms: give up (and set ctrl.timed_out)
// instead of hanging when a peer never arrives.
CEDNERF_EXPORT int cednerf_dp_barrier(const CednerfDpPeers* peers, uint32_t epoch, int timeout_ms, void* stream) {
  CEDNERF_REQUIRE(peers && peers->world >= 1 && peers->world <= DP_MAX_RANKS && peers->rank >= 0 && peers->rank < peers->world,
                  "bad peer table");
  if (peers->world == 1) return 0;
  dp_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*peers, epoch, (uint64_t)timeout_ms * 1000000ull);
  return cednerf_check_launch("cednerf_dp_barrier");
}

// After the gradient barrier: found_out (device float, nullable) = any rank found inf / nan; *step (nullable) advances
// unless so.
CEDNERF_EXPORT int cednerf_dp_found_inf(const CednerfDpPeers* peers, float* found_out, float* step, void* stream) {
  CEDNERF_REQUIRE(peers && peers->world >= 1 && peers->world <= DP_MAX_RANKS, "bad peer table");
  dp_found_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(*peers, found_out, step);
  return cednerf_check_launch("cednerf_dp_found_inf");
}

// Fused reduce-scatter + unscale + Adam + all-gather of the owned range (see the file header).  `step` is the shared
// device step counter AFTER cednerf_dp_found_inf advanced it; grad_scale / found_inf as in cednerf_adam_step.
CEDNERF_EXPORT int cednerf_dp_adam(const CednerfDpAdam* args, const float* step, const float* grad_scale,
                                   const float* found_inf, float beta1, float beta2, float eps, int adam_w_mode,
                                   void* stream) {
  CEDNERF_REQUIRE(args && step, "bad arguments");
  const CednerfDpAdam& a = *args;
  CEDNERF_REQUIRE(a.world >= 1 && a.world <= DP_MAX_RANKS && a.n_out >= 1 && a.n_out <= DP_MAX_RANKS, "bad world / n_out");
  CEDNERF_REQUIRE(a.lo >= 0 && a.hi >= a.lo && (a.lo & 3) == 0 && a.m && a.v && a.p32_out[0], "bad range or state");
  for (int r = 0; r < a.world; ++r) CEDNERF_REQUIRE(a.grad[r] && (((uintptr_t)a.grad[r]) & 15) == 0, "gradient buffers: 16-byte aligned");
  for (int r = 0; r < a.n_out; ++r)
    CEDNERF_REQUIRE(a.p32_out[r] && (((uintptr_t)a.p32_out[r]) & 15) == 0 && (((uintptr_t)a.p16_out[r]) & 7) == 0,
                    "parameter replicas: 16-byte (fp32) / 8-byte (fp16) aligned");
  CEDNERF_REQUIRE(((((uintptr_t)a.m) | ((uintptr_t)a.v)) & 15) == 0, "moments: 16-byte aligned");
  const int64_t n = a.hi - a.lo;
  if (n == 0) return 0;
  const int64_t ctas = (n + DP_CHUNK - 1) / DP_CHUNK;
  dp_adam_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(a, step, grad_scale, found_inf, beta1, beta2, eps,
                                                                   adam_w_mode);
  return cednerf_check_launch("cednerf_dp_adam");
}
