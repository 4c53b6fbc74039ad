// K2 — multi-resolution hash-grid encode (3-D, and 4-D xyz+t key-frame variant) and backward.
//
// Same algorithm as the spatial encoder the reference executes through tcnn.Encoding(HashGrid)
// (cednerf/model.py:242-252, :384) and spells out in-tree in
//   cednerf/taichi_kernel/hash_encoder_half.py:66-103 (index / scale), :112-161 (fwd), :164-226 (bwd)
//   cednerf/taichi_kernel/hash_encoder_inter.py:121-199 (4-D fwd), :202-275 (4-D bwd)
// One thread per (sample, level), level fastest, so a sample's 2*L features are written as one
// contiguous run and its xyz is a warp broadcast.  All 8 corner gathers are issued before use.
// Numerics shared with oracle/tcnn_ref.py: pos = x*scale (+) 0.5 as two rounded fp32 ops, fp32
// trilinear weights in corner-bit order, fp16 table, fp32 accumulate, one rounding to fp16 on output.
#include "hashgrid.cuh"

namespace {

__global__ void __launch_bounds__(256)
hashgrid_fwd_kernel(const float* __restrict__ x, int x_stride, int64_t n, const __half* __restrict__ table,
                    CednerfGridLevels lv, __half* __restrict__ out, int out_stride) {
  const int L = lv.n_levels;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = tid / L;
  const int l = (int)(tid - s * L);
  if (s >= n) return;
  const float* xs = x + s * x_stride;
  const Cell c = locate(xs, lv.scale[l]);
  const uint32_t res = lv.res[l], size = lv.size[l], off = lv.offset[l];
  const bool hashed = lv.hashed[l] != 0;
  uint32_t idx[8];
  float wgt[8];
  cell_corners(c, res, size, off, hashed, idx, wgt);
  float a0 = 0.f, a1 = 0.f;
  {
    __half2 v[8];
    const __half2* t2 = reinterpret_cast<const __half2*>(table);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(t2 + idx[k]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 f = __half22float2(v[k]);
      a0 = __fadd_rn(a0, __fmul_rn(wgt[k], f.x));
      a1 = __fadd_rn(a1, __fmul_rn(wgt[k], f.y));
    }
  }
  *reinterpret_cast<__half2*>(out + s * out_stride + 2 * l) = __floats2half2_rn(a0, a1);
}

// 4-D forward, one thread per SAMPLE walking the levels (the stand-alone 3-D forward above keeps a thread per (sample,
// level)): the lanes of a warp are 32 consecutive samples - neighbours on a ray - so a level's eight 16-byte gathers
// touch one or two sectors per warp at the coarse and middle levels instead of 32 (with level-fastest lanes every lane of
// a gather sat in a different level, i.e. a different sector: the L1 data stage ran at 66 % with DRAM at 52 %), and a
// sample's 32 features leave as four 16-byte stores.  Two levels (16 gathers of 16 bytes) in flight per thread.  Same
// arithmetic per (sample, level), operation by operation, as the kernel above: bit-identical features.
template <int LG>
__global__ void __launch_bounds__(256)
hashgrid4d_fwd_sample_kernel(const float* __restrict__ x, int x_stride, int64_t n, const __half* __restrict__ table,
                             CednerfGridLevels lv, __half* __restrict__ out, int out_stride, int taichi_compat) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const float xs[4] = {x[s * x_stride], x[s * x_stride + 1], x[s * x_stride + 2], x[s * x_stride + 3]};
  int kf;
  float tau;
  keyframe(xs[3], taichi_compat, kf, tau);
  const float om = __fsub_rn(1.f, tau);
  const uint4* t8 = reinterpret_cast<const uint4*>(table);
  const int L = lv.n_levels;
  uint32_t* orow = reinterpret_cast<uint32_t*>(out + s * out_stride);
  uint32_t pend[4];
#pragma unroll 1
  for (int l0 = 0; l0 < L; l0 += LG) {
    uint4 v[LG][8];
    float wgt[LG][8];
#pragma unroll
    for (int a = 0; a < LG; ++a) {
      const int l = l0 + a < L ? l0 + a : L - 1;
      const Cell c = locate(xs, lv.scale[l]);
      uint32_t idx[8];
      cell_corners(c, lv.res[l], lv.size[l], lv.offset[l], lv.hashed[l] != 0, idx, wgt[a]);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[a][k] = __ldg(t8 + idx[k]);
    }
#pragma unroll
    for (int a = 0; a < LG; ++a) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t lo_w = kf == 0 ? v[a][k].x : (kf == 1 ? v[a][k].y : v[a][k].z);
        const uint32_t hi_w = kf == 0 ? v[a][k].y : (kf == 1 ? v[a][k].z : v[a][k].w);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&lo_w));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&hi_w));
        const float w = wgt[a][k];
        a0 = __fadd_rn(a0, __fmul_rn(w, __fadd_rn(__fmul_rn(lo.x, om), __fmul_rn(hi.x, tau))));
        a1 = __fadd_rn(a1, __fmul_rn(w, __fadd_rn(__fmul_rn(lo.y, om), __fmul_rn(hi.y, tau))));
      }
      const __half2 h = __floats2half2_rn(a0, a1);
      const int l = l0 + a;
      if (l < L) {
        pend[l & 3] = *reinterpret_cast<const uint32_t*>(&h);
        const bool row_aligned = ((out_stride & 7) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        if ((l & 3) == 3 && row_aligned) {
          *reinterpret_cast<uint4*>(orow + (l - 3)) = make_uint4(pend[0], pend[1], pend[2], pend[3]);
        } else if (!row_aligned || (l == L - 1 && (l & 3) != 3)) {
          if (!row_aligned) orow[l] = pend[l & 3];
          else
            for (int q = l & ~3; q <= l; ++q) orow[q] = pend[q & 3];
        }
      }
    }
  }
}

// Backward: table gradient (fp32 vector reductions, red.global.add.v2.f32) and, for the 3-D encoder,
// the input gradient dL/dx_d = scale_l * sum_corners (+-) prod_{e != d} w_e * <table[idx], dy_l>
// (tcnn's kernel_grid_backward_input form; the in-tree Taichi form w/(+-f) is 0/0 on cell faces).
template <bool FOUR_D, typename GradT>
__global__ void __launch_bounds__(256)
hashgrid_bwd_kernel(const float* __restrict__ x, int x_stride, int64_t n, const __half* __restrict__ table,
                    CednerfGridLevels lv, const GradT* __restrict__ dy, int dy_stride, float* __restrict__ g_table,
                    float* __restrict__ g_x, int taichi_compat) {
  const int L = lv.n_levels;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = tid / L;
  const int l = (int)(tid - s * L);
  const bool active = s < n;
  if (!active) s = n - 1;  // keep the lane alive for the shuffles below
  const float* xs = x + s * x_stride;
  const Cell c = locate(xs, lv.scale[l]);
  const uint32_t res = lv.res[l], size = lv.size[l], off = lv.offset[l];
  const bool hashed = lv.hashed[l] != 0;
  float d0 = 0.f, d1 = 0.f;
  if (active) {
    d0 = (float)dy[s * dy_stride + 2 * l];
    d1 = (float)dy[s * dy_stride + 2 * l + 1];
  }
  uint32_t idx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    idx[k] = off + corner_index(c.g[0] + (k & 1), c.g[1] + ((k >> 1) & 1), c.g[2] + ((k >> 2) & 1), res, size, hashed);

  float gx[3] = {0.f, 0.f, 0.f};
  if (g_x) {
    __half2 v[8];
    const __half2* t2 = reinterpret_cast<const __half2*>(table);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(t2 + idx[k]);
    float dot[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 f = __half22float2(v[k]);
      dot[k] = f.x * d0 + f.y * d1;
    }
    const float fx = c.f[0], fy = c.f[1], fz = c.f[2];
    const float wx[2] = {1.f - fx, fx}, wy[2] = {1.f - fy, fy}, wz[2] = {1.f - fz, fz};
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        gx[0] += wy[a] * wz[b] * (dot[1 + 2 * a + 4 * b] - dot[0 + 2 * a + 4 * b]);
        gx[1] += wx[a] * wz[b] * (dot[a + 2 + 4 * b] - dot[a + 0 + 4 * b]);
        gx[2] += wx[a] * wy[b] * (dot[a + 2 * b + 4] - dot[a + 2 * b]);
      }
    const float sc = lv.scale[l];
    gx[0] *= sc;
    gx[1] *= sc;
    gx[2] *= sc;
    if (!active) gx[0] = gx[1] = gx[2] = 0.f;
  }
  if (g_table && active && (d0 != 0.f || d1 != 0.f)) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float w = corner_weight(c, k);
      atomicAdd(reinterpret_cast<float2*>(g_table) + idx[k], make_float2(w * d0, w * d1));
    }
  }
  if (g_x) {
    if (L == 8 || L == 16 || L == 32) {  // the L lanes of a sample are contiguous in the warp
      for (int o = L / 2; o > 0; o >>= 1) {
        gx[0] += __shfl_xor_sync(0xffffffffu, gx[0], o);
        gx[1] += __shfl_xor_sync(0xffffffffu, gx[1], o);
        gx[2] += __shfl_xor_sync(0xffffffffu, gx[2], o);
      }
      if (active && l == 0) {
        g_x[3 * s] = gx[0];
        g_x[3 * s + 1] = gx[1];
        g_x[3 * s + 2] = gx[2];
      }
    } else if (active) {  // caller zero-fills g_x
      atomicAdd(g_x + 3 * s, gx[0]);
      atomicAdd(g_x + 3 * s + 1, gx[1]);
      atomicAdd(g_x + 3 * s + 2, gx[2]);
    }
  }
}

// Table gradient, level-major: blockIdx.y = level, lanes = 32 CONSECUTIVE samples.  Samples are packed along rays, so
// at coarse and middle levels neighbouring lanes sit in the same cell and would hammer the same 8 addresses (on the
// DyNeRF scene the content occupies [-1,1]^3 of a [-8,8]^3 grid: the 16^3 level sees ~64 distinct entries in total).
// Lanes are grouped into runs of equal cell; each run is summed with a segmented shuffle scan and only its last lane
// issues the vector reduction.  Fine levels degenerate to runs of one lane (one reduction per corner, as before).
//
// Dense (coarse) levels additionally go through a CTA-local accumulation cache (CACHED): every sample of the batch
// lands in the same few hundred entries there (the 16^3 level of the DyNeRF-shaped scene sees ~30 distinct entries for
// 8 M corner updates), and L2 serialises reductions per address - the six dense levels took 1.1 ms of the 1.3 ms this
// kernel needed.  A CTA walks a chunk of several thousand samples, adds into a direct-mapped shared-memory table
// (slot = index mod TG_SLOTS, claimed with a compare-and-swap on its tag; a slot held by another entry falls back to
// the global reduction) and flushes each occupied slot once at the end: reductions per hot address drop by the chunk
// length over the warp size.
#define TG_SLOTS 2048
#define TG_EMPTY 0xffffffffu

struct LevelList {
  int n;
  int id[CEDNERF_MAX_LEVELS];
};

// FOUR_D (xyz + t key-frame table, 8 floats per entry: 4 key-frames x 2 features, hash_encoder_inter.py:202-275): a corner
// receives (w dy om, w dy tau) at key-frames k and k + 1, i.e. FOUR CONSECUTIVE floats at float offset 2k of its entry - one
// 16-byte reduction when k is even, two 8-byte ones when k = 1.  Runs additionally break where the key-frame changes
// (samples of one ray share their timestamp, so in practice they do not).
template <typename GradT, bool LEVEL_MAJOR, bool CACHED, bool FOUR_D = false>
__global__ void __launch_bounds__(256)
hashgrid_bwd_table_kernel(const float* __restrict__ x, int x_stride, int64_t n, CednerfGridLevels lv, LevelList list,
                          const GradT* __restrict__ dy, int dy_stride, float* __restrict__ g_table, int64_t chunk,
                          const int64_t* __restrict__ n_dev, int taichi_compat = 0) {
  constexpr int EW = FOUR_D ? 8 : 2;                          // floats per table entry
  constexpr int SLOTS = FOUR_D ? TG_SLOTS / 2 : TG_SLOTS;     // 4-D: 1024 slots x 32 B = 32 KB of shared memory
  __shared__ uint32_t tags[CACHED ? SLOTS : 1];
  __shared__ float vals[CACHED ? EW * SLOTS : 1];
  // n: capacity of the sample arrays (and the level stride of a level-major dy); n_dev (nullable): live sample count
  int64_t n_live = n;
  if (n_dev) {
    const int64_t v = *n_dev;
    n_live = v < n ? v : n;
  }
  if ((int64_t)blockIdx.x * chunk >= n_live) return;
  const int l = list.id[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const uint32_t res = lv.res[l], size = lv.size[l], off = lv.offset[l];
  const bool hashed = lv.hashed[l] != 0;
  const float scale = lv.scale[l];
  float2* t2 = reinterpret_cast<float2*>(g_table) + off;
  if (CACHED) {
    for (int q = threadIdx.x; q < SLOTS; q += blockDim.x) tags[q] = TG_EMPTY;
    for (int q = threadIdx.x; q < EW * SLOTS; q += blockDim.x) vals[q] = 0.f;
    __syncthreads();
  }
  // 4-D: four floats at float offset 2 kf of entry i
  auto add4 = [&](uint32_t i, int kf, const float* v) {
    if (v[0] == 0.f && v[1] == 0.f && v[2] == 0.f && v[3] == 0.f) return;
    if (CACHED) {
      const uint32_t slot = i & (SLOTS - 1);
      const uint32_t old = atomicCAS(&tags[slot], TG_EMPTY, i);
      if (old == TG_EMPTY || old == i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&vals[EW * slot + 2 * kf + j], v[j]);
        return;
      }
    }
    float* dst = g_table + ((size_t)off + i) * 8 + 2 * kf;
    if ((kf & 1) == 0) {
      atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
    } else {
      atomicAdd(reinterpret_cast<float2*>(dst), make_float2(v[0], v[1]));
      atomicAdd(reinterpret_cast<float2*>(dst) + 1, make_float2(v[2], v[3]));
    }
  };
  auto add = [&](uint32_t i, float a, float b) {  // i: index inside the level
    if (a == 0.f && b == 0.f) return;
    if (CACHED) {
      const uint32_t slot = i & (SLOTS - 1);
      const uint32_t old = atomicCAS(&tags[slot], TG_EMPTY, i);
      if (old == TG_EMPTY || old == i) {
        atomicAdd(&vals[2 * slot], a);
        atomicAdd(&vals[2 * slot + 1], b);
        return;
      }
    }
    atomicAdd(t2 + i, make_float2(a, b));
  };
  const int64_t begin = (int64_t)blockIdx.x * chunk, end = begin + chunk < n_live ? begin + chunk : n_live;
  for (int64_t s0 = begin; s0 < end; s0 += blockDim.x) {
    int64_t s = s0 + threadIdx.x;
    const bool active = s < end;
    if (!active) s = end - 1;
    const Cell c = locate(x + s * x_stride, scale);
    float d0 = 0.f, d1 = 0.f;
    if (active) {
      if (LEVEL_MAJOR) {  // dy stored [level][sample][2]: one coalesced 4-byte read per lane, each byte read once
        d0 = (float)dy[((int64_t)l * n + s) * 2];
        d1 = (float)dy[((int64_t)l * n + s) * 2 + 1];
      } else {
        d0 = (float)dy[s * dy_stride + 2 * l];
        d1 = (float)dy[s * dy_stride + 2 * l + 1];
      }
    }
    // run structure: a lane starts a run when its cell differs from the previous lane's
    const uint32_t px = __shfl_up_sync(0xffffffffu, c.g[0], 1), py = __shfl_up_sync(0xffffffffu, c.g[1], 1),
                   pz = __shfl_up_sync(0xffffffffu, c.g[2], 1);
    int kf = 0;
    float tau = 0.f, om = 1.f;
    if (FOUR_D) {
      keyframe(x[s * x_stride + 3], taichi_compat, kf, tau);
      om = 1.f - tau;
    }
    const int pk = __shfl_up_sync(0xffffffffu, kf, 1);
    const bool head = lane == 0 || px != c.g[0] || py != c.g[1] || pz != c.g[2] || (FOUR_D && pk != kf);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int run_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // highest head at or below this lane
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    // scan depth follows the longest run in the warp: 0 steps at fine levels (every lane its own run), 5 at the coarsest
    const int max_len = __reduce_max_sync(0xffffffffu, lane - run_start + 1);
    // corners come in x-pairs (k, k+1): when the two table entries form an aligned 16-byte pair (dense levels: even
    // index; hashed levels: even x, because the x term of the hash is x itself) one 16-byte reduction carries both.
    // Weights and indices with shared sub-expressions (cell_weights; hashed: per-axis terms masked once, one xor per
    // corner; dense: base + offsets unless a corner wraps) - same values as corner_weight / corner_index.
    float w8[8];
    cell_weights(c.f, w8);
    uint32_t hy[2], hz[2], x0, x1;
    bool plain;  // indices are (x0 | x1) combined with hy / hz by xor (hashed) or by addition (dense, no wrap)
    if (hashed) {
      const uint32_t mask = size - 1u;
      const uint32_t y0 = c.g[1] * 2654435761u, z0 = c.g[2] * 805459861u;
      x0 = c.g[0] & mask, x1 = (c.g[0] + 1u) & mask;
      hy[0] = y0 & mask, hy[1] = (y0 + 2654435761u) & mask;
      hz[0] = z0 & mask, hz[1] = (z0 + 805459861u) & mask;
      plain = true;
    } else {
      const uint32_t r2 = res * res;
      x0 = c.g[0] + c.g[1] * res + c.g[2] * r2, x1 = x0 + 1u;
      hy[0] = 0u, hy[1] = res, hz[0] = 0u, hz[1] = r2;
      plain = x0 < size - (1u + res + r2);
    }
    if constexpr (FOUR_D) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v[4];
        v[0] = w8[k] * d0 * om, v[1] = w8[k] * d1 * om, v[2] = w8[k] * d0 * tau, v[3] = w8[k] * d1 * tau;
        for (int o = 1; o < max_len; o <<= 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float u = __shfl_up_sync(0xffffffffu, v[j], o);
            if (lane - o >= run_start) v[j] += u;
          }
        }
        if (tail) {
          uint32_t i;
          if (hashed) i = ((k & 1) ? x1 : x0) ^ hy[(k >> 1) & 1] ^ hz[k >> 2];
          else if (plain) i = ((k & 1) ? x1 : x0) + hy[(k >> 1) & 1] + hz[k >> 2];
          else i = corner_index(c.g[0] + (k & 1), c.g[1] + ((k >> 1) & 1), c.g[2] + (k >> 2), res, size, false);
          add4(i, kf, v);
        }
      }
    } else {
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      float v[4];
      v[0] = w8[2 * kp] * d0, v[1] = w8[2 * kp] * d1, v[2] = w8[2 * kp + 1] * d0, v[3] = w8[2 * kp + 1] * d1;
      for (int o = 1; o < max_len; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float u = __shfl_up_sync(0xffffffffu, v[j], o);
          if (lane - o >= run_start) v[j] += u;
        }
      }
      if (tail) {
        uint32_t i0, i1;
        if (hashed) {
          const uint32_t yz = hy[kp & 1] ^ hz[kp >> 1];
          i0 = x0 ^ yz, i1 = x1 ^ yz;
        } else if (plain) {
          const uint32_t yz = hy[kp & 1] + hz[kp >> 1];
          i0 = x0 + yz, i1 = x1 + yz;
        } else {
          const uint32_t gy = c.g[1] + (kp & 1), gz = c.g[2] + (kp >> 1);
          i0 = corner_index(c.g[0], gy, gz, res, size, false);
          i1 = corner_index(c.g[0] + 1, gy, gz, res, size, false);
        }
        if (!CACHED && (i0 ^ i1) == 1u && ((off & 1u) == 0u)) {
          const bool even = (i0 & 1u) == 0u;
          const float4 val = even ? make_float4(v[0], v[1], v[2], v[3]) : make_float4(v[2], v[3], v[0], v[1]);
          if (val.x != 0.f || val.y != 0.f || val.z != 0.f || val.w != 0.f)
            atomicAdd(reinterpret_cast<float4*>(t2 + (i0 & ~1u)), val);
        } else {
          add(i0, v[0], v[1]);
          add(i1, v[2], v[3]);
        }
      }
    }
    }  // 3-D
  }
  if (CACHED) {
    __syncthreads();
    for (int q = threadIdx.x; q < SLOTS; q += blockDim.x) {
      const uint32_t i = tags[q];
      if (i == TG_EMPTY) continue;
      if (FOUR_D) {
        float4* dst = reinterpret_cast<float4*>(g_table + ((size_t)off + i) * 8);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 val = make_float4(vals[EW * q + 4 * h], vals[EW * q + 4 * h + 1], vals[EW * q + 4 * h + 2],
                                         vals[EW * q + 4 * h + 3]);
          if (val.x != 0.f || val.y != 0.f || val.z != 0.f || val.w != 0.f) atomicAdd(dst + h, val);
        }
      } else {
        const float a = vals[2 * q], b = vals[2 * q + 1];
        if (a != 0.f || b != 0.f) atomicAdd(t2 + i, make_float2(a, b));
      }
    }
  }
}

// small dense levels -> cached pass (long chunks), the others -> direct pass (one block of 256 samples per CTA)
template <typename GradT, bool LEVEL_MAJOR, bool FOUR_D = false>
int launch_table_gradient(const float* x, int x_stride, int64_t n, const CednerfGridLevels& lv, const GradT* dy, int dy_stride,
                          float* g_table, cudaStream_t st, const int64_t* n_dev = nullptr, int taichi_compat = 0) {
  LevelList cached{}, direct{};
  for (int l = 0; l < lv.n_levels; ++l) {
    // the cache pays where a level has few entries (heavy per-address contention in L2); from 2^20 entries on - hashed
    // levels and the finest dense one - the direct reductions already run at the L2 reduction rate (measured per level)
    LevelList& dst = (lv.hashed[l] || lv.size[l] >= (1u << 20)) ? direct : cached;
    dst.id[dst.n++] = l;
  }
  int launches = 0;
  if (cached.n) {
    const int64_t want_ctas = (int64_t)cednerf_num_sms() * 8 / cached.n + 1;
    int64_t chunk = ((n + want_ctas - 1) / want_ctas + 255) / 256 * 256;
    if (chunk < 2048) chunk = 2048;
    dim3 grid((unsigned)((n + chunk - 1) / chunk), cached.n);
    hashgrid_bwd_table_kernel<GradT, LEVEL_MAJOR, true, FOUR_D><<<grid, 256, 0, st>>>(x, x_stride, n, lv, cached, dy, dy_stride,
                                                                                 g_table, chunk, n_dev, taichi_compat);
    ++launches;
  }
  if (direct.n) {
    dim3 grid(cednerf_blocks(n, 256), direct.n);
    hashgrid_bwd_table_kernel<GradT, LEVEL_MAJOR, false, FOUR_D><<<grid, 256, 0, st>>>(x, x_stride, n, lv, direct, dy, dy_stride,
                                                                                  g_table, 256, n_dev, taichi_compat);
    ++launches;
  }
  return launches;
}

__global__ void cast_f32_to_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4,
                                       const float* __restrict__ src_tail, __half* __restrict__ dst_tail, int tail) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = src[i];
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    dst[i] = o;
  }
  if (i < tail) dst_tail[i] = __float2half_rn(src_tail[i]);
}

int check_levels(const CednerfGridLevels* lv) {
  if (!lv || lv->n_levels < 1 || lv->n_levels > CEDNERF_MAX_LEVELS) return 0;
  for (int l = 0; l < lv->n_levels; ++l) {
    if (lv->size[l] == 0) return 0;
    // hashed levels are reduced with `& (size - 1)` in every kernel (forward, table gradient, fused field kernels)
    if (lv->hashed[l] && (lv->size[l] & (lv->size[l] - 1u)) != 0u) return 0;
  }
  return 1;
}

}  // namespace

CEDNERF_EXPORT int cednerf_hashgrid_fwd(const float* x, int x_stride, int64_t n, const void* table_f16,
                                        const CednerfGridLevels* levels, void* out_f16, int out_stride,
                                        void* stream) {
  CEDNERF_REQUIRE(check_levels(levels), "bad level table");
  CEDNERF_REQUIRE(n >= 0 && x_stride >= 3 && out_stride >= 2 * levels->n_levels && (out_stride % 2) == 0, "bad sizes");
  if (n == 0) return 0;
  hashgrid_fwd_kernel<<<cednerf_blocks(n * levels->n_levels, 256), 256, 0, (cudaStream_t)stream>>>(
      x, x_stride, n, (const __half*)table_f16, *levels, (__half*)out_f16, out_stride);
  return cednerf_check_launch("cednerf_hashgrid_fwd");
}

CEDNERF_EXPORT int cednerf_hashgrid_bwd(const float* x, int x_stride, int64_t n, const void* table_f16,
                                        const CednerfGridLevels* levels, const void* dy, int dy_stride,
                                        int dy_is_f16, float* g_table, float* g_x, void* stream) {
  CEDNERF_REQUIRE(check_levels(levels), "bad level table");
  CEDNERF_REQUIRE(n >= 0 && x_stride >= 3 && dy_stride >= 2 * levels->n_levels, "bad sizes");
  CEDNERF_REQUIRE(g_table || g_x, "nothing to compute");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  if (g_table) {
    if (dy_is_f16)
      launches += launch_table_gradient<__half, false>(x, x_stride, n, *levels, (const __half*)dy, dy_stride, g_table, st) - 1;
    else
      launches += launch_table_gradient<float, false>(x, x_stride, n, *levels, (const float*)dy, dy_stride, g_table, st) - 1;
    ++launches;
  }
  if (g_x) {
    dim3 grid(cednerf_blocks(n * levels->n_levels, 256));
    if (dy_is_f16)
      hashgrid_bwd_kernel<false, __half><<<grid, 256, 0, st>>>(x, x_stride, n, (const __half*)table_f16, *levels,
                                                             (const __half*)dy, dy_stride, nullptr, g_x, 0);
    else
      hashgrid_bwd_kernel<false, float><<<grid, 256, 0, st>>>(x, x_stride, n, (const __half*)table_f16, *levels,
                                                            (const float*)dy, dy_stride, nullptr, g_x, 0);
    ++launches;
  }
  return cednerf_check_launch("cednerf_hashgrid_bwd", launches);
}

// Table gradient from a LEVEL-MAJOR fp16 gradient dy_lm[level][sample][2] (what the fused density-net backward
// writes): every level pass streams its own 4 bytes/sample instead of striding through a [n, 2L] row-major matrix.
CEDNERF_EXPORT int cednerf_hashgrid_bwd_table_lm(const float* x, int x_stride, int64_t n, const CednerfGridLevels* levels,
                                                 const void* dy_lm_f16, float* g_table, void* stream) {
  CEDNERF_REQUIRE(check_levels(levels), "bad level table");
  CEDNERF_REQUIRE(n >= 0 && x_stride >= 3 && dy_lm_f16 && g_table, "bad arguments");
  if (n == 0) return 0;
  const int launches = launch_table_gradient<__half, true>(x, x_stride, n, *levels, (const __half*)dy_lm_f16, 0, g_table,
                                                           (cudaStream_t)stream);
  return cednerf_check_launch("cednerf_hashgrid_bwd_table_lm", launches);
}

// the same with the live sample count on the device (n = capacity = level stride of dy_lm); used by the fused backward
extern "C" int cednerf_hashgrid_bwd_table_lm_dev(const float* x, int x_stride, int64_t n, const int64_t* n_device,
                                                 const CednerfGridLevels* levels, const void* dy_lm_f16, float* g_table,
                                                 void* stream) {
  CEDNERF_REQUIRE(check_levels(levels), "bad level table");
  CEDNERF_REQUIRE(n >= 0 && x_stride >= 3 && dy_lm_f16 && g_table, "bad arguments");
  if (n == 0) return 0;
  const int launches = launch_table_gradient<__half, true>(x, x_stride, n, *levels, (const __half*)dy_lm_f16, 0, g_table,
                                                           (cudaStream_t)stream, n_device);
  return cednerf_check_launch("cednerf_hashgrid_bwd_table_lm", launches);
}

CEDNERF_EXPORT int cednerf_hashgrid4d_fwd(const float* xyzt, int x_stride, int64_t n, const void* table_f16,
                                          const CednerfGridLevels* levels, void* out_f16, int out_stride,
                                          int taichi_compat, void* stream) {
  CEDNERF_REQUIRE(check_levels(levels), "bad level table");
  CEDNERF_REQUIRE(n >= 0 && x_stride >= 4 && out_stride >= 2 * levels->n_levels && (out_stride % 2) == 0, "bad sizes");
  CEDNERF_REQUIRE(((uintptr_t)table_f16 & 15) == 0, "4-D table must be 16-byte aligned");
  if (n == 0) return 0;
  // two levels in flight per thread (107 registers, 16 warps / SM): 1.93 ms on 2^22 ray-coherent samples against 2.02 ms
  // with one level (56 registers) and 2.66 ms with the thread-per-(sample, level) kernel (profiles/r2q_hash4d_full_size.md)
  hashgrid4d_fwd_sample_kernel<2><<<cednerf_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(
      xyzt, x_stride, n, (const __half*)table_f16, *levels, (__half*)out_f16, out_stride, taichi_compat);
  return cednerf_check_launch("cednerf_hashgrid4d_fwd");
}

CEDNERF_EXPORT int cednerf_hashgrid4d_bwd(const float* xyzt, int x_stride, int64_t n, const CednerfGridLevels* levels,
                                          const void* dy, int dy_stride, int dy_is_f16, float* g_table,
                                          int taichi_compat, void* stream) {
  CEDNERF_REQUIRE(check_levels(levels), "bad level table");
  CEDNERF_REQUIRE(n >= 0 && x_stride >= 4 && dy_stride >= 2 * levels->n_levels && g_table, "bad sizes");
  if (n == 0) return 0;
  CEDNERF_REQUIRE(((uintptr_t)g_table & 15) == 0, "4-D gradient table must be 16-byte aligned");
  // level-major walk with run merging and the shared-memory cache for the small dense levels, as the 3-D table gradient
  const int launches =
      dy_is_f16 ? launch_table_gradient<__half, false, true>(xyzt, x_stride, n, *levels, (const __half*)dy, dy_stride, g_table,
                                                             (cudaStream_t)stream, nullptr, taichi_compat)
                : launch_table_gradient<float, false, true>(xyzt, x_stride, n, *levels, (const float*)dy, dy_stride, g_table,
                                                            (cudaStream_t)stream, nullptr, taichi_compat);
  return cednerf_check_launch("cednerf_hashgrid4d_bwd", launches);
}

// fp32 master parameters -> fp16 working copy (the reference does this cast on every forward,
// hash_encoder_half.py:381-385; here it runs once per parameter version)
CEDNERF_EXPORT int cednerf_cast_f32_to_f16(const float* src, void* dst_f16, int64_t n, void* stream) {
  CEDNERF_REQUIRE(n >= 0, "bad size");
  if (n == 0) return 0;
  const int64_t n4 = n / 4;
  const int tail = (int)(n - n4 * 4);
  CEDNERF_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst_f16 & 7) == 0, "unaligned buffers");
  cast_f32_to_f16_kernel<<<cednerf_blocks(n4 > tail ? n4 : tail, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)src, (uint2*)dst_f16, n4, src + n4 * 4, (__half*)dst_f16 + n4 * 4, tail);
  return cednerf_check_launch("cednerf_cast_f32_to_f16");
}
