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

#include "tc05.cuh"

#define DP_MAX_RANKS 8
#define DP_CHUNK 4096  // elements per CTA of the direct-load kernel
#define DP_TILE 2048   // elements per pipeline stage and peer of the TMA-staged kernel (8 KB)
#define DP_STAGES 3
#define DP_TMA_THREADS 512

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
  float* p32_out[DP_MAX_RANKS];     // fp32 parameter replicas: entry 0 = the local master (required), others nullable
  void* p16_out[DP_MAX_RANKS];      // fp16 working copies, nullable
  float* m;                         // Adam moments of the OWNED range, indexed from lo
  float* v;
  int64_t lo, hi;                   // owned element range; lo is a multiple of 4
  float lr, weight_decay, grad_div; // the summed gradient is divided by grad_div (world for an average)
  // NVLink SHARP (NVLS), when the buffers are multicast-mapped (torch symmetric memory): grad_mc = the multicast address of
  // the gradient buffers - one multimem.ld_reduce returns the sum over all ranks, formed inside the switch; p16_mc = the
  // multicast address of the fp16 working copies - one multimem.st updates every replica.  Both nullable.
  const float* grad_mc;
  void* p16_mc;
};

// The small (MLP) parameter tensors of a step in ONE launch: every rank sums every peer's copy of the staged gradients in
// rank order (bit-identical replicas without a broadcast) and applies Adam to its own replica.
#define DP_SMALL_MAX 8
#define DP_SMALL_CHUNK 1024
struct CednerfDpSmall {
  int world, n_tensors;
  const float* grad[DP_MAX_RANKS];   // every rank's staging region of the small gradients (same layout), as mapped here
  float* p[DP_SMALL_MAX];            // this rank's parameters, moments
  float* m[DP_SMALL_MAX];
  float* v[DP_SMALL_MAX];
  int64_t off[DP_SMALL_MAX];         // element offset of the tensor's gradient inside the staging region
  int64_t n[DP_SMALL_MAX];
  float lr[DP_SMALL_MAX];
  float weight_decay[DP_SMALL_MAX];
  float grad_div;
  int64_t chunk_begin[DP_SMALL_MAX + 1];  // scratch, filled by the library
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

// streaming (evict-first) 16-byte load; deliberately NOT volatile so that the compiler may hoist all loads of a chunk
// ahead of the arithmetic (remote loads take microseconds: the more of them in flight, the closer to link bandwidth)
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

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
    constexpr int IT = DP_CHUNK / 4 / 256;
    int64_t e[IT];
    float4 g[IT], pp[IT], mm[IT], vv[IT];
    // phase 1: every load of this thread's IT float4 items is issued before anything is consumed
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      e[j] = base + (int64_t)(j * 256 + threadIdx.x) * 4;
      g[j] = ld_stream4(a.grad[0] + e[j]);
    }
#pragma unroll
    for (int r = 1; r < DP_MAX_RANKS; ++r)
      if (r < a.world) {
        float4 x[IT];
#pragma unroll
        for (int j = 0; j < IT; ++j) x[j] = ld_stream4(a.grad[r] + e[j]);
#pragma unroll
        for (int j = 0; j < IT; ++j) g[j].x += x[j].x, g[j].y += x[j].y, g[j].z += x[j].z, g[j].w += x[j].w;
      }
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      pp[j] = *reinterpret_cast<const float4*>(p_in + e[j]);
      mm[j] = ld_stream4(a.m + (e[j] - a.lo));
      vv[j] = ld_stream4(a.v + (e[j] - a.lo));
    }
    // phase 2: update; phase 3: local state, then the replicas
#pragma unroll
    for (int j = 0; j < IT; ++j) {
      adam_one(pp[j].x, g[j].x, mm[j].x, vv[j].x, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp[j].y, g[j].y, mm[j].y, vv[j].y, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp[j].z, g[j].z, mm[j].z, vv[j].z, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      adam_one(pp[j].w, g[j].w, mm[j].w, vv[j].w, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
      __stcs(reinterpret_cast<float4*>(a.m + (e[j] - a.lo)), mm[j]);
      __stcs(reinterpret_cast<float4*>(a.v + (e[j] - a.lo)), vv[j]);
    }
#pragma unroll
    for (int r = 0; r < DP_MAX_RANKS; ++r)
      if (r < a.n_out) {
#pragma unroll
        for (int j = 0; j < IT; ++j) {
          if (a.p32_out[r]) *reinterpret_cast<float4*>(a.p32_out[r] + e[j]) = pp[j];
          if (a.p16_out[r]) {
            const __half2 h0 = __floats2half2_rn(pp[j].x, pp[j].y), h1 = __floats2half2_rn(pp[j].z, pp[j].w);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&h0);
            o.y = *reinterpret_cast<const uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.p16_out[r]) + e[j]) = o;
          }
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
        if (a.p32_out[r]) a.p32_out[r][e] = pp;
        if (a.p16_out[r]) reinterpret_cast<__half*>(a.p16_out[r])[e] = __float2half_rn(pp);
      }
    }
  }
  __threadfence_system();  // the replica stores are performed before this rank's next barrier signal
}

__global__ void __launch_bounds__(256) dp_adam_small_kernel(CednerfDpSmall a, const float* step, const float* scale_p,
                                                            const float* found_inf, float b1, float b2, float eps, int adamw) {
  if (found_inf && *found_inf != 0.f) return;
  int k = 0;
  while (k + 1 < a.n_tensors && (int64_t)blockIdx.x >= a.chunk_begin[k + 1]) ++k;
  const int64_t base = ((int64_t)blockIdx.x - a.chunk_begin[k]) * DP_SMALL_CHUNK;
  const float inv_scale = scale_p ? 1.f / *scale_p : 1.f;
  const float g_mul = a.grad_div != 1.f ? inv_scale / a.grad_div : inv_scale;
  const float st = *step;
  const float bc1 = 1.f - powf(b1, st), bc2_sqrt = sqrtf(1.f - powf(b2, st));
  for (int64_t i = base + threadIdx.x; i < a.n[k] && i < base + DP_SMALL_CHUNK; i += 256) {
    float x[DP_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < DP_MAX_RANKS; ++r) x[r] = r < a.world ? __ldcs(a.grad[r] + a.off[k] + i) : 0.f;  // all in flight
    float g = x[0];
#pragma unroll
    for (int r = 1; r < DP_MAX_RANKS; ++r)
      if (r < a.world) g += x[r];
    float pp = a.p[k][i], mm = a.m[k][i], vv = a.v[k][i];
    adam_one(pp, g, mm, vv, g_mul, a.lr[k], a.weight_decay[k], adamw, b1, b2, eps, bc1, bc2_sqrt);
    a.p[k][i] = pp, a.m[k][i] = mm, a.v[k][i] = vv;
  }
}

// NVLS variant: the reduce-scatter is ONE instruction per 16 bytes (multimem.ld_reduce: the NVSwitch fetches the element
// from every rank's buffer and returns the fp32 sum, so only the owned slice crosses this GPU's ingress links instead of
// N - 1 copies of it) and the all-gather of the fp16 copy is one multimem.st (one copy leaves the GPU, the switch fans it
// out).  Everything else - unscale, Adam on the owned slice, local master and moments - is as in the kernels below.
__device__ __forceinline__ float4 multimem_ld_reduce_add4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_h4(void* mc, uint32_t lo, uint32_t hi) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f16x2 [%0], {%1, %2};" ::"l"(mc), "r"(lo), "r"(hi) : "memory");
}

__global__ void __launch_bounds__(256) dp_adam_nvls_kernel(CednerfDpAdam a, const float* step, const float* scale_p,
                                                           const float* found_inf, float b1, float b2, float eps, int adamw) {
  if (found_inf && *found_inf != 0.f) return;
  const int64_t base = a.lo + (int64_t)blockIdx.x * DP_CHUNK;   // the launcher passes full chunks only
  const float inv_scale = scale_p ? 1.f / *scale_p : 1.f;
  const float g_mul = a.grad_div != 1.f ? inv_scale / a.grad_div : inv_scale;
  const float st = *step;
  const float bc1 = 1.f - powf(b1, st), bc2_sqrt = sqrtf(1.f - powf(b2, st));
  float* p_loc = a.p32_out[0];
  constexpr int IT = DP_CHUNK / 4 / 256;
  int64_t e[IT];
  float4 g[IT], pp[IT], mm[IT], vv[IT];
#pragma unroll
  for (int j = 0; j < IT; ++j) {
    e[j] = base + (int64_t)(j * 256 + threadIdx.x) * 4;
    g[j] = multimem_ld_reduce_add4(a.grad_mc + e[j]);
  }
#pragma unroll
  for (int j = 0; j < IT; ++j) {
    pp[j] = *reinterpret_cast<const float4*>(p_loc + e[j]);
    mm[j] = ld_stream4(a.m + (e[j] - a.lo));
    vv[j] = ld_stream4(a.v + (e[j] - a.lo));
  }
#pragma unroll
  for (int j = 0; j < IT; ++j) {
    adam_one(pp[j].x, g[j].x, mm[j].x, vv[j].x, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    adam_one(pp[j].y, g[j].y, mm[j].y, vv[j].y, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    adam_one(pp[j].z, g[j].z, mm[j].z, vv[j].z, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    adam_one(pp[j].w, g[j].w, mm[j].w, vv[j].w, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    __stcs(reinterpret_cast<float4*>(a.m + (e[j] - a.lo)), mm[j]);
    __stcs(reinterpret_cast<float4*>(a.v + (e[j] - a.lo)), vv[j]);
    *reinterpret_cast<float4*>(p_loc + e[j]) = pp[j];
    const __half2 h0 = __floats2half2_rn(pp[j].x, pp[j].y), h1 = __floats2half2_rn(pp[j].z, pp[j].w);
    multimem_st_h4(reinterpret_cast<__half*>(a.p16_mc) + e[j], *reinterpret_cast<const uint32_t*>(&h0),
                   *reinterpret_cast<const uint32_t*>(&h1));
  }
  __threadfence_system();
}

// The same step for large ranges, with the peers' gradient tiles STAGED THROUGH SHARED MEMORY BY TMA: a persistent CTA
// walks tiles of DP_TILE elements; one elected thread keeps DP_STAGES tiles ahead with cp.async.bulk (one 8 KB bulk copy
// per peer and tile, completion on an mbarrier), so the microsecond NVLink reads never occupy registers or stall the
// threads that stream the local state (own gradient, p, m, v) at HBM speed.  With direct loads the remote fetch, the
// local traffic and the replica stores of a CTA ran one after the other (measured on 2 GPUs: 0.47 ms = 0.17 local +
// 0.13 remote in + 0.17 out; the three overlap here).  Sum order: own gradient first, then the peers by rank.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(DP_TMA_THREADS) dp_adam_tma_kernel(CednerfDpAdam a, int64_t n_tiles, const float* step,
                                                                      const float* scale_p, const float* found_inf, float b1,
                                                                      float b2, float eps, int adamw) {
  if (found_inf && *found_inf != 0.f) return;
  extern __shared__ __align__(128) uint8_t dp_smem[];
  __shared__ uint64_t full[DP_STAGES];
  const int n_rem = a.world - 1;
  float* stage_base = reinterpret_cast<float*>(dp_smem);  // [DP_STAGES][n_rem][DP_TILE] peers' gradient tiles
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < DP_STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int64_t tile, int s) {  // thread 0: the peers' gradient tiles of `tile` -> stage s
    mbar_expect_tx(&full[s], (uint32_t)(n_rem * DP_TILE * 4));
    int k = 0;
    for (int r = 0; r < a.world; ++r) {
      if (r == a.rank) continue;
      bulk_g2s(stage_base + ((size_t)s * n_rem + k) * DP_TILE, a.grad[r] + a.lo + tile * DP_TILE, DP_TILE * 4, &full[s]);
      ++k;
    }
  };
  if (tid == 0)
    for (int s = 0; s < DP_STAGES; ++s) {
      const int64_t t = blockIdx.x + (int64_t)s * gridDim.x;
      if (t < n_tiles) issue(t, s);
    }
  const float inv_scale = scale_p ? 1.f / *scale_p : 1.f;
  const float g_mul = a.grad_div != 1.f ? inv_scale / a.grad_div : inv_scale;
  const float st = *step;
  const float bc1 = 1.f - powf(b1, st), bc2_sqrt = sqrtf(1.f - powf(b2, st));
  const float* g_own = a.grad[a.rank];
  const float* p_in = a.p32_out[0];
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % DP_STAGES;
    const int64_t e = a.lo + tile * DP_TILE + (int64_t)tid * 4;   // DP_TILE == 4 * DP_TMA_THREADS: one float4 per thread
    // local state first: these loads are in flight while we wait for the staged tile
    float4 g = ld_stream4(g_own + e);
    float4 pp = *reinterpret_cast<const float4*>(p_in + e);
    float4 mm = ld_stream4(a.m + (e - a.lo));
    float4 vv = ld_stream4(a.v + (e - a.lo));
    mbar_wait(&full[s], (uint32_t)((it / DP_STAGES) & 1));
    const float4* staged = reinterpret_cast<const float4*>(stage_base + (size_t)s * n_rem * DP_TILE) + tid;
    for (int k = 0; k < n_rem; ++k) {
      const float4 x = staged[(size_t)k * (DP_TILE / 4)];
      g.x += x.x, g.y += x.y, g.z += x.z, g.w += x.w;
    }
    __syncthreads();  // every thread has read stage s: it can be refilled
    if (tid == 0) {
      const int64_t nxt = tile + (int64_t)DP_STAGES * gridDim.x;
      if (nxt < n_tiles) issue(nxt, s);
    }
    adam_one(pp.x, g.x, mm.x, vv.x, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    adam_one(pp.y, g.y, mm.y, vv.y, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    adam_one(pp.z, g.z, mm.z, vv.z, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    adam_one(pp.w, g.w, mm.w, vv.w, g_mul, a.lr, a.weight_decay, adamw, b1, b2, eps, bc1, bc2_sqrt);
    __stcs(reinterpret_cast<float4*>(a.m + (e - a.lo)), mm);
    __stcs(reinterpret_cast<float4*>(a.v + (e - a.lo)), vv);
    const __half2 h0 = __floats2half2_rn(pp.x, pp.y), h1 = __floats2half2_rn(pp.z, pp.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h0);
    o.y = *reinterpret_cast<const uint32_t*>(&h1);
#pragma unroll
    for (int r = 0; r < DP_MAX_RANKS; ++r)   // replicas: entry 0 is local; remote fp32 entries are optional.  (Staging the
      if (r < a.n_out) {                     // fp16 tile in shared memory for TMA bulk stores was measured 8 % slower.)
        if (a.p32_out[r]) *reinterpret_cast<float4*>(a.p32_out[r] + e) = pp;
        if (a.p16_out[r]) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(a.p16_out[r]) + e) = o;
      }
  }
  __threadfence_system();
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
    CEDNERF_REQUIRE((r > 0 || a.p32_out[r]) && (((uintptr_t)a.p32_out[r]) & 15) == 0 && (((uintptr_t)a.p16_out[r]) & 7) == 0,
                    "parameter replicas: 16-byte (fp32) / 8-byte (fp16) aligned; entry 0 is the local fp32 master");
  CEDNERF_REQUIRE(((((uintptr_t)a.m) | ((uintptr_t)a.v)) & 15) == 0, "moments: 16-byte aligned");
  static_assert(DP_TILE == 4 * DP_TMA_THREADS, "one float4 per thread and tile");
  const int64_t n = a.hi - a.lo;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  CednerfDpAdam rest = a;
  // multicast-mapped buffers: reduce-scatter and all-gather inside the NVSwitch (full chunks; the tail goes direct)
  if (a.world > 1 && a.grad_mc && a.p16_mc && n >= DP_CHUNK && (((uintptr_t)a.grad_mc | (uintptr_t)a.p16_mc) & 15) == 0) {
    const int64_t chunks = n / DP_CHUNK;
    dp_adam_nvls_kernel<<<(unsigned)chunks, 256, 0, st>>>(a, step, grad_scale, found_inf, beta1, beta2, eps, adam_w_mode);
    ++launches;
    rest.lo = a.lo + chunks * DP_CHUNK;
    rest.m = a.m + chunks * DP_CHUNK;
    rest.v = a.v + chunks * DP_CHUNK;
    if (rest.hi > rest.lo) {
      const int64_t ctas = (rest.hi - rest.lo + DP_CHUNK - 1) / DP_CHUNK;
      dp_adam_kernel<<<(unsigned)ctas, 256, 0, st>>>(rest, step, grad_scale, found_inf, beta1, beta2, eps, adam_w_mode);
      ++launches;
    }
    return cednerf_check_launch("cednerf_dp_adam", launches);
  }
  // large ranges with remote gradients: TMA-staged persistent kernel on the full tiles, direct-load kernel on the tail
  const int64_t n_tiles = n / DP_TILE;
  const int smem = DP_STAGES * (a.world - 1) * DP_TILE * 4;
  if (a.world > 1 && n_tiles >= 64 && a.rank >= 0 && a.rank < a.world && smem <= 200 * 1024) {
    static CednerfOncePerDevice configured;
    if (int e = cednerf_opt_in_smem(dp_adam_tma_kernel, 200 * 1024, configured, "cednerf_dp_adam")) return e;
    int per_sm = 1;  // persistent CTAs: exactly as many as are resident at once
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_adam_tma_kernel, DP_TMA_THREADS, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    int64_t ctas = (int64_t)cednerf_num_sms() * per_sm;
    if (ctas > n_tiles) ctas = n_tiles;
    dp_adam_tma_kernel<<<(unsigned)ctas, DP_TMA_THREADS, smem, st>>>(a, n_tiles, step, grad_scale, found_inf, beta1, beta2,
                                                                     eps, adam_w_mode);
    ++launches;
    rest.lo = a.lo + n_tiles * DP_TILE;
    rest.m = a.m + n_tiles * DP_TILE;
    rest.v = a.v + n_tiles * DP_TILE;
  }
  if (rest.hi > rest.lo) {
    const int64_t ctas = (rest.hi - rest.lo + DP_CHUNK - 1) / DP_CHUNK;
    dp_adam_kernel<<<(unsigned)ctas, 256, 0, st>>>(rest, step, grad_scale, found_inf, beta1, beta2, eps, adam_w_mode);
    ++launches;
  }
  return cednerf_check_launch("cednerf_dp_adam", launches);
}

// Adam on up to 8 small replicated tensors in one launch; gradients = rank-ordered sum of every rank's staging region.
CEDNERF_EXPORT int cednerf_dp_adam_small(const CednerfDpSmall* args, const float* step, const float* grad_scale,
                                         const float* found_inf, float beta1, float beta2, float eps, int adam_w_mode,
                                         void* stream) {
  CEDNERF_REQUIRE(args && step && args->world >= 1 && args->world <= DP_MAX_RANKS && args->n_tensors >= 0 &&
                      args->n_tensors <= DP_SMALL_MAX,
                  "bad arguments");
  CednerfDpSmall a = *args;
  int64_t c = 0;
  for (int k = 0; k < a.n_tensors; ++k) {
    CEDNERF_REQUIRE(a.n[k] >= 0 && (a.n[k] == 0 || (a.p[k] && a.m[k] && a.v[k])), "bad tensor entry");
    a.chunk_begin[k] = c;
    c += (a.n[k] + DP_SMALL_CHUNK - 1) / DP_SMALL_CHUNK;
  }
  a.chunk_begin[a.n_tensors] = c;
  for (int r = 0; r < a.world; ++r) CEDNERF_REQUIRE(a.grad[r], "gradient staging regions missing");
  if (c == 0) return 0;
  dp_adam_small_kernel<<<(unsigned)c, 256, 0, (cudaStream_t)stream>>>(a, step, grad_scale, found_inf, beta1, beta2, eps,
                                                                      adam_w_mode);
  return cednerf_check_launch("cednerf_dp_adam_small");
}
