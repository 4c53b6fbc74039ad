// tcgen05 / TMEM / mbarrier PTX wrappers and the 128-byte-swizzle tile helpers shared by the MLP and field kernels.
#pragma once
#include "common.cuh"

#define MLP_MAX_LAYERS 5
#define MLP_TILE 128
#define MLP_ROW_BYTES 128
#define MLP_TILE_BYTES (MLP_TILE * MLP_ROW_BYTES)

struct CednerfMlpDesc {
  int n_layers;                    // hidden layers + 1
  int dim_in[MLP_MAX_LAYERS];      // padded (multiple of 16, <= 64)
  int dim_out[MLP_MAX_LAYERS];     // 64 for hidden layers, padded n_out (16..64) for the last
  int param_off[MLP_MAX_LAYERS];   // element offset of W_l ([dim_out, dim_in] row-major) in the flat fp32 params
  int image_off[MLP_MAX_LAYERS];   // byte offset of W_l's swizzled fp16 image (dim_out rows x 128 B)
  int image_bytes;
};

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait suspends the thread in hardware for up to the hinted time, so a waiting warp costs a few issue slots per
  // wait instead of a tight polling loop (the loop was 13 % of the field kernel's executed instructions)
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
    if (++spins > (1u << 22)) __trap();  // a lost completion must surface as an error, never as a hang
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_enc, uint32_t sbo_enc) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)lbo_enc << 16) | ((uint64_t)sbo_enc << 32) | (1ull << 46) |
         (2ull << 61);
}
// instruction descriptor for kind::f16: fp16 A/B, fp32 D
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 columns x this warp's 32 lanes <- one value (used to zero accumulators that several issuers add into)
__device__ __forceinline__ void tmem_st16_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of 16-byte chunk `c` of row `r` inside a swizzled [rows][128 B] tile (tile base 1024-aligned)
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * MLP_ROW_BYTES + ((c ^ (r & 7)) << 4)); }

// global [rows_valid, width] fp16 row-major  ->  swizzled tile; rows >= rows_valid are zero-filled
__device__ __forceinline__ void load_tile(uint8_t* tile, const __half* __restrict__ g, int width, int rows_valid) {
  const int cpr = width >> 3;  // 16-byte chunks per row
  const int total = MLP_TILE * cpr;
  const uint4* src = reinterpret_cast<const uint4*>(g);
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int r = q / cpr, c = q - r * cpr;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid) v = __ldg(src + q);
    *reinterpret_cast<uint4*>(tile + swz(r, c)) = v;
  }
}
// same as load_tile but with cp.async (LDGSTS): the copies are in flight while the thread goes on to wait for the tensor
// core and run its epilogue; cp_async_wait_all() + fence.proxy.async must precede the MMA that reads the tile
__device__ __forceinline__ void load_tile_async(uint8_t* tile, const __half* __restrict__ g, int width, int rows_valid,
                                                int tid = -1, int n_threads = 0) {
  const int cpr = width >> 3;
  const int total = MLP_TILE * cpr;
  const uint4* src = reinterpret_cast<const uint4*>(g);
  if (tid < 0) tid = threadIdx.x, n_threads = blockDim.x;  // default: the whole CTA loads the tile
  for (int q = tid; q < total; q += n_threads) {
    const int r = q / cpr, c = q - r * cpr;
    const uint32_t dst = smem_u32(tile + swz(r, c));
    const int bytes = r < rows_valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + (r < rows_valid ? q : 0)), "r"(bytes)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// swizzled tile -> global [rows_valid, width] fp16 row-major (coalesced 16-byte chunks)
__device__ __forceinline__ void store_tile(const uint8_t* tile, __half* __restrict__ g, int width, int rows_valid) {
  const int cpr = width >> 3;
  const int total = rows_valid * cpr;
  uint4* dst = reinterpret_cast<uint4*>(g);
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int r = q / cpr, c = q - r * cpr;
    dst[q] = *reinterpret_cast<const uint4*>(tile + swz(r, c));
  }
}

// ReLU fused into the fp32 -> fp16x2 conversion (cvt.rn.relu): one instruction per pair of activations
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}


}  // namespace
