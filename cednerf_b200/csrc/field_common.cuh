// Shared pieces of the fused field kernels (inference: field.cu, training: field_train.cu).
#pragma once
#include "hashgrid.cuh"
#include "tc05.cuh"

struct CednerfFieldDesc {
  float aabb[6];
  float moving_step;
  int use_div_offsets;    // deformation net emits 6 values: move = o[:3]*MS + tanh(o[3:])*MS   (model.py:358-363)
  int time_mode;          // 0 none, 1 SinusoidalEncoder, 2 SinusoidalEncoderWithExp (attenuated by |move|)
  int time_before_sigma;  // 1: density-MLP input = [hash | time9];  0: colour-MLP input = [sh4 | feat15 | time9]
  CednerfMlpDesc f1, f2, f3;  // deformation, density, colour networks
  CednerfMlpDesc f4;          // hash-feature predictor (-f, training only); n_layers == 0 when absent
  CednerfGridLevels levels;
};

namespace {

__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

// Layers of one network on the 128-row tile in `abuf` (one warp-group of 4 warps = one tile).  Hidden activations are
// written back IN PLACE: the MMA that read the tile has completed (commit -> mbarrier) before any row is overwritten.
// The last layer's accumulator is left in TMEM columns [0, N_last) of this group's TMEM slice.
__device__ __forceinline__ void run_chain(const CednerfMlpDesc& d, const uint8_t* wimg, uint8_t* abuf, uint32_t tmem_grp,
                                          uint32_t tmem_warp, uint64_t* bar, uint32_t& phase, int gtid, int group,
                                          uint8_t* save = nullptr, int64_t save_layer_stride = 0, int64_t save_row0 = 0,
                                          int save_rows = 0, uint8_t* save_in = nullptr) {

  // `save` (training forward): the post-ReLU activation tiles are kept for the backward pass AS THEY SIT IN SHARED
  // MEMORY - the 16 KB swizzled image of tile t, layer l at save + (l * save_layer_stride + t) * 16384 with
  // save_layer_stride = number of tiles and t = save_row0 / 128 - so that one elected thread stores a tile with one TMA
  // bulk copy (cp.async.bulk.global.shared::cta, UBLKCP) and the backward fetches it with one bulk load straight into
  // its operand buffer: no per-thread address arithmetic, no 8 x (LDS + STG) per thread and layer.  The buffer is
  // private to the library (cednerf_field_saved_bytes), so its layout is ours to choose.
  const int L = d.n_layers;
  for (int l = 0; l < L; ++l) {
    const int K = d.dim_in[l], N = d.dim_out[l];
    if (gtid == 0) {
      tc_fence_after();
      const uint64_t ad = make_desc(smem_u32(abuf), 1, 64);
      const uint64_t bd = make_desc(smem_u32(wimg + d.image_off[l]), 1, 64);
      const uint32_t id = make_idesc(128, N, 0, 0);
      for (int k = 0; k < K / 16; ++k) umma(tmem_grp, ad + 2 * k, bd + 2 * k, id, k > 0);
      umma_commit(bar);
    }
    if (save_in && l == 0) {  // the network's input tile ([rows, K] fp16), copied out the same cooperative way
      const int cpr = K >> 3;
      uint4* dst = reinterpret_cast<uint4*>(save_in + save_row0 * K * 2);
      for (int q = gtid; q < MLP_TILE * cpr; q += MLP_TILE) {
        const int r = q / cpr, c = q - r * cpr;
        if (r < save_rows) dst[q] = *reinterpret_cast<const uint4*>(abuf + swz(r, c));
      }
    }
    if (save && l > 0 && gtid == 0) {
      // the tile this layer's MMA is reading (post-ReLU activations of the previous layer) leaves by TMA while the MMA
      // runs; the store has finished READING shared memory before anyone overwrites the tile (wait_group.read below)
      uint8_t* dst = save + ((int64_t)(l - 1) * save_layer_stride + (save_row0 >> 7)) * MLP_TILE_BYTES;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(abuf)),
                   "r"((uint32_t)MLP_TILE_BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // one warp of the group polls the mbarrier; the other three block on the group's named barrier, which costs no
    // issue slots (four polling warps per tile were 14 % of the kernel's executed instructions)
    if ((gtid >> 5) == 0) {
      mbar_wait(bar, phase);
      tc_fence_after();
      tc_fence_before();
      if (save && l > 0 && gtid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    group_sync(group);
    phase ^= 1;
    tc_fence_after();
    if (l < L - 1) {
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t r[16];
        tmem_ld16(tmem_warp + cb * 16, r);
        tmem_ld_wait();
        uint32_t p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = pack_relu_h2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        *reinterpret_cast<uint4*>(abuf + swz(gtid, 2 * cb)) = make_uint4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<uint4*>(abuf + swz(gtid, 2 * cb + 1)) = make_uint4(p[4], p[5], p[6], p[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      group_sync(group);
    }
  }
}

// value the tensor-core chain hands on: fp32 accumulator rounded to fp16 (tcnn network output precision)
__device__ __forceinline__ float rnd16(uint32_t acc_bits) { return __half2float(__float2half_rn(__uint_as_float(acc_bits))); }

// Time embedding on the special-function unit: arguments are t 2^i (+ pi/2) with t in [0, 1], i <= 3, i.e. below 10 rad,
// where __sinf's fp32 argument scaling is good to ~6e-7 absolute; the nine values are rounded to fp16 (half ulp 2.4e-4)
// as soon as they enter the MLP input row.  (Full-precision sinf was 1200 SASS instructions and ~160 executed per sample.)
// EXACT = true (gradient-carrying kernels): sinf / expf, the arithmetic of the stand-alone encoder (encodings.cu).
template <bool EXACT = false>
__device__ __forceinline__ void time_embedding(float tv, float mvnorm, int mode, float* e /*[9]*/) {
  const float half_pi = 1.5707963267948966f;
  e[0] = tv;
  if (mode == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float tb = tv * (float)(1 << i);
      e[1 + i] = EXACT ? sinf(tb) : __sinf(tb);
      e[5 + i] = EXACT ? sinf(tb + half_pi) : __sinf(tb + half_pi);
    }
  } else {
    const float scm[4] = {0.f, 2.f, 8.f, 24.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float tb = tv * (float)(1 << i);
      const float att = EXACT ? expf(-1.f * (mvnorm * scm[i])) : __expf(-1.f * (mvnorm * scm[i]));
      e[1 + 2 * i] = (EXACT ? sinf(tb) : __sinf(tb)) * att;
      e[2 + 2 * i] = (EXACT ? sinf(tb + half_pi) : __sinf(tb + half_pi)) * att;
    }
  }
}

__device__ __noinline__ uint32_t wrapped_corner(uint32_t base, int k, uint32_t res, uint32_t r2, uint32_t size) {
  const uint32_t h = base + (uint32_t)(k & 1) + (uint32_t)((k >> 1) & 1) * res + (uint32_t)(k >> 2) * r2;
  return h >= size ? h % size : h;
}

// table[idx] with the address formed by ONE 64-bit multiply-add (the compiler otherwise folds the level offset into
// every corner's index and spends five instructions per address on 64-bit carries)
__device__ __forceinline__ __half2 ldg_last(const __half2* p) {  // as gather_h2, from a formed address
  uint32_t v;
  asm("ld.global.nc.L1::evict_last.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return *reinterpret_cast<__half2*>(&v);
}
__device__ __forceinline__ __half2 gather_h2(const __half2* base, uint32_t idx) {
  uint64_t addr;
  asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(addr) : "r"(idx), "l"(base));
  // L1::evict_last: the gathered lines are the only reusable data these kernels touch (x-neighbour corners, neighbouring
  // samples, the coarse levels of every sample); measured against the default policy: field_fwd 1.393 -> 1.382 ms,
  // training forward 0.565 -> 0.544 ms.  no_allocate: 3.05 ms; evict_first on the fine levels only: 2.23 ms.
  uint32_t v;
  asm("ld.global.nc.L1::evict_last.b32 %0, [%1];" : "=r"(v) : "l"(addr));
  return *reinterpret_cast<__half2*>(&v);
}

// packed fp32 pairs (sm_100: add / sub / mul / fma .f32x2 on 64-bit registers)
__device__ __forceinline__ uint64_t pack_f2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t sub_f2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t h2_to_f2(__half2 h) {
  const float2 f = __half22float2(h);
  return pack_f2(f.x, f.y);
}

#ifndef HASH_PAIR_LOADS
#define HASH_PAIR_LOADS 1       // measured: field_fwd 1.387 -> 1.340 ms, training forward 0.543 -> 0.525 ms
#endif
#ifndef HASH_PAIR_MIN_RES
#define HASH_PAIR_MIN_RES 1024u // levels finer than this (11-15 of the DyNeRF table)
#endif
// the 8*LG gathers of levels l_first .. l_first+LG-1 (fractions kept for the weights)
template <int LG>
__device__ __forceinline__ void hash_issue(const float* xn, const __half* __restrict__ table, const CednerfGridLevels& lv,
                                           int l_first, float (*frac)[3], __half2 (*v)[8]) {
#pragma unroll
  for (int a = 0; a < LG; ++a) {
    const int l = l_first + a;
    const Cell c = locate(xn, lv.scale[l]);
    frac[a][0] = c.f[0], frac[a][1] = c.f[1], frac[a][2] = c.f[2];
    const uint32_t res = lv.res[l], size = lv.size[l];
    const __half2* tl;  // this level's table, as one opaque 64-bit value
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(tl) : "r"(lv.offset[l]), "l"(table));
    if (lv.hashed[l]) {  // power-of-two table: (x ^ y P1 ^ z P2) & mask == (x & mask) ^ (y P1 & mask) ^ (z P2 & mask)
      const uint32_t mask = size - 1u;
      const uint32_t y0 = c.g[1] * 2654435761u, z0 = c.g[2] * 805459861u;
      const uint32_t hx[2] = {c.g[0] & mask, (c.g[0] + 1u) & mask};
      const uint32_t hy[2] = {y0 & mask, (y0 + 2654435761u) & mask};
      const uint32_t hz[2] = {z0 & mask, (z0 + 805459861u) & mask};
      if (HASH_PAIR_LOADS && res > HASH_PAIR_MIN_RES && (reinterpret_cast<uintptr_t>(tl) & 7u) == 0u) {
        // the finest levels: every lane sits in its own sector, and the L1 tag stage paces the kernel.  The x term of the
        // hash is x itself, so for even x the two x-neighbours of a corner pair are the two halves of one aligned 8-byte
        // pair: one look-up instead of two for half of the lanes (odd x: two 4-byte gathers, as before).  Same values.
        const bool even = (c.g[0] & 1u) == 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t yz = hy[j & 1] ^ hz[j >> 1];
          const uint32_t A = hx[0] ^ yz;
          if (even) {
            uint64_t addr;
            asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(addr) : "r"(A & ~1u), "l"(tl));
            uint32_t lo, hi;
            asm("ld.global.nc.L1::evict_last.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(addr));
            const bool odd = (A & 1u) != 0u;
            const uint32_t va = odd ? hi : lo, vb = odd ? lo : hi;
            v[a][2 * j] = *reinterpret_cast<const __half2*>(&va);
            v[a][2 * j + 1] = *reinterpret_cast<const __half2*>(&vb);
          } else {
            v[a][2 * j] = gather_h2(tl, A);
            v[a][2 * j + 1] = gather_h2(tl, hx[1] ^ yz);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[a][k] = gather_h2(tl, hx[k & 1] ^ hy[(k >> 1) & 1] ^ hz[k >> 2]);
      }
    } else {
      const uint32_t r2 = res * res;
      const uint32_t base = c.g[0] + c.g[1] * res + c.g[2] * r2;
      if (base < size - (1u + res + r2)) {  // all eight corners in range: no wrap
        const __half2* p = tl + base;
        v[a][0] = ldg_last(p), v[a][1] = ldg_last(p + 1);
        v[a][2] = ldg_last(p + res), v[a][3] = ldg_last(p + res + 1);
        v[a][4] = ldg_last(p + r2), v[a][5] = ldg_last(p + r2 + 1);
        v[a][6] = ldg_last(p + r2 + res), v[a][7] = ldg_last(p + r2 + res + 1);
      } else {  // points outside the box (masked by the selector afterwards): the reference's wrap, out of line - it
                // is rare, and eight inlined integer modulos per level were 600 SASS instructions
#pragma unroll
        for (int k = 0; k < 8; ++k) v[a][k] = __ldg(tl + wrapped_corner(base, k, res, r2, size));
      }
    }
  }
}

template <int LG>
__device__ __forceinline__ void hash_levels(const float* xn, const __half* __restrict__ table, const CednerfGridLevels& lv,
                                            int l0, uint32_t* feat, int level_base = 0) {
  float frac[LG][3];
  __half2 v[LG][8];
  hash_issue<LG>(xn, table, lv, level_base + l0, frac, v);  // issue all 8*LG gathers first ...
#pragma unroll
  for (int a = 0; a < LG; ++a) {
    // ... then the blend, as seven lerps on both features at once (Blackwell's packed fp32 pipe: FADD2 / FFMA2): 14
    // instructions instead of 12 weight products + 16 FMAs.  Same value as the weighted sum up to fp32 rounding, i.e.
    // the same fp16 feature except at near-ties (the stand-alone encoder keeps the oracle's operation order).
    const uint64_t fx = pack_f2(frac[a][0], frac[a][0]), fy = pack_f2(frac[a][1], frac[a][1]),
                   fz = pack_f2(frac[a][2], frac[a][2]);
    uint64_t c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t lo = h2_to_f2(v[a][2 * j]), hi = h2_to_f2(v[a][2 * j + 1]);
      c[j] = fma_f2(fx, sub_f2(hi, lo), lo);
    }
    const uint64_t d0 = fma_f2(fy, sub_f2(c[1], c[0]), c[0]), d1 = fma_f2(fy, sub_f2(c[3], c[2]), c[2]);
    const uint64_t e = fma_f2(fz, sub_f2(d1, d0), d0);
    float a0, a1;
    unpack_f2(e, a0, a1);
    feat[l0 + a] = pack_h2(a0, a1);
  }
}

// sin(pi y), cos(pi y) on the special-function unit: y - 2 rint(y / 2) is exact in fp32 and lies in [-1, 1], where
// MUFU.SIN / MUFU.COS are accurate to 2^-21 absolute (tcnn's Frequency encoding uses __sinf the same way); the result is
// rounded to fp16 (half ulp 2^-12) right after.
__device__ __forceinline__ void sincospi_fast(float y, float& sn, float& cs) {
  const float r = fmaf(-2.f, rintf(0.5f * y), y);
  const float a = r * 3.14159265358979323846f;
  sn = __sinf(a);
  cs = __cosf(a);
}

// position / time of packed sample s (cednerf/utils.py:74-104): x = o + (d * (t0 + t1)) / 2, individually rounded
__device__ __forceinline__ void packed_sample(const int64_t* __restrict__ ridx, const float* __restrict__ t0,
                                              const float* __restrict__ t1, const float* __restrict__ rays_o,
                                              const float* __restrict__ rays_d, const float* __restrict__ ts, int t_stride,
                                              int64_t s, float* x, float& tv, int64_t& ray) {
  ray = ridx[s];
  const float tm = __fadd_rn(t0[s], t1[s]);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    x[k] = __fadd_rn(rays_o[3 * ray + k], __fmul_rn(__fmul_rn(rays_d[3 * ray + k], tm), 0.5f));
  tv = ts[ray * t_stride];
}

// Frequency(4 dims, 4 octaves) of (v0, v1, v2, v3) -> 32 halves in chunks 0..3 of row `row` of a swizzled tile.
// EXACT = false: MUFU.SIN / MUFU.COS (2^-21 ABSOLUTE error: values near a zero crossing can land one fp16 ulp off) - the
// no-grad kernels (sampler pre-pass, eval rendering).  EXACT = true: sincospif (1 ulp of fp32, i.e. the correctly rounded
// fp16 value except at near-ties) - the gradient-carrying kernels: a hidden unit of the deformation net whose
// pre-activation sits near zero has its ReLU mask decided by that last fp16 ulp of its inputs, and the mask then
// multiplies every sample's gradient (measured on the D-NeRF-shaped configuration, 2.8 M samples: one such unit put the
// deformation net's weight gradient 4.6e-3 away from the oracle's; with exact inputs it is 6e-4, profiles/r2*_parity*).
template <bool EXACT = false>
__device__ __forceinline__ void frequency_row(uint8_t* tile, int row, float v0, float v1, float v2, float v3) {
  const float in4[4] = {v0, v1, v2, v3};
#pragma unroll
  for (int dim = 0; dim < 4; ++dim) {
    uint32_t p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sn, cs;  // sin(pi ph) and sin(pi (ph + 1/2)) = cos(pi ph): one range reduction for the pair
      if (EXACT) sincospif(in4[dim] * (float)(1 << k), &sn, &cs);
      else sincospi_fast(in4[dim] * (float)(1 << k), sn, cs);
      p[k] = pack_h2(sn, cs);
    }
    *reinterpret_cast<uint4*>(tile + swz(row, dim)) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

// deformation-net output (fp16 accumulator values) -> move, normalised position, selector (model.py:354-383)
__device__ __forceinline__ void apply_move(const CednerfFieldDesc& d, const float* x, const float* off /*[6]*/, float* mv,
                                           float* xn, bool& selector) {
  selector = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mv[k] = off[k] * d.moving_step;
    if (d.use_div_offsets) mv[k] = mv[k] + tanhf(off[3 + k]) * d.moving_step;
    const float xm = x[k] + mv[k];
    xn[k] = __fdiv_rn(__fsub_rn(xm, d.aabb[k]), __fsub_rn(d.aabb[3 + k], d.aabb[k]));
    selector = selector && (xn[k] > 0.f) && (xn[k] < 1.f);
  }
}

// colour-net input row [SH4(dir) | feat15 | (time 9) | 1.0 ...] -> 32 halves, chunks 0..3 (model.py:447-466)
__device__ __forceinline__ void colour_input_row(const CednerfFieldDesc& d, uint8_t* tile, int row, const float* dir,
                                                 const float* feat15, const float* temb) {
  const float nrm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  float v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = ((dir[k] / nrm + 1.f) / 2.f) * 2.f - 1.f;
  float in[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) in[j] = 1.f;
  in[0] = 0.28209479177387814f;
  in[1] = -0.48860251190291987f * v[1];
  in[2] = 0.48860251190291987f * v[2];
  in[3] = -0.48860251190291987f * v[0];
#pragma unroll
  for (int j = 0; j < 15; ++j) in[4 + j] = feat15[j];
  if (d.time_mode && !d.time_before_sigma) {
#pragma unroll
    for (int j = 0; j < 9; ++j) in[19 + j] = temb[j];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + swz(row, c)) =
        make_uint4(pack_h2(in[8 * c], in[8 * c + 1]), pack_h2(in[8 * c + 2], in[8 * c + 3]),
                   pack_h2(in[8 * c + 4], in[8 * c + 5]), pack_h2(in[8 * c + 6], in[8 * c + 7]));
}

}  // namespace
