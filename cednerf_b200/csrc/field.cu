// Fused radiance-field query: positions -> deformation MLP -> hash-grid encode -> density MLP (-> colour MLP)
// in ONE kernel per 128-sample tile (no activation, feature or position tensor ever touches HBM).
//
// Replaces the chain DNGPradianceField.query_density / .forward executes as ~40 separate PyTorch / tcnn launches
// (cednerf/model.py:354-365 query_move, :378-384 normalise + selector + hash, :386-403 time embedding, :406-417
// density MLP + trunc_exp, :447-466 SH + colour MLP + sigmoid), together with the position closure of
// cednerf/utils.py:74-104 (x = o + d * (t0 + t1) / 2, t = timestamps[ray]).
//
// Per tile: thread r owns sample r end to end (its position, Frequency / SH / time encodings, its 16x8 hash
// gathers, its activations after each layer); the only cross-thread step is the tensor-core layer itself:
// operand rows are written to a 128-byte-swizzled shared-memory tile, one thread issues tcgen05.mma against the
// weight images resident in shared memory, the accumulator comes back from TMEM with tcgen05.ld.
// Adjacent threads hold adjacent samples of the same ray, so one gather instruction touches one level for 32
// neighbouring samples: coarse and middle levels coalesce into few sectors, L1/L2 serve the reuse.
#include "field_common.cuh"

namespace {

struct FieldFwdArgs {
  const int64_t* ridx;  // packed samples (with t0, t1, rays_o, rays_d) ...
  const float* t0;
  const float* t1;
  const float* rays_o;
  const float* rays_d;
  const float* x;       // ... or explicit points (with dirs)
  const float* dirs;
  const float* t;       // timestamps: [n_rays] when ridx != null else [n]; stride 0 = one value for all
  int t_stride;
  int64_t n;
  const int64_t* n_dev;  // nullable: the live sample count on the device (n is then the capacity of the buffers)
  const uint8_t* img1;
  const uint8_t* img2;
  const uint8_t* img3;
  const __half* table;
  float* sigma;   // [n] (nullable in cell mode)
  float* rgb;     // [n,3] or null (density only)
  // cell mode (occupancy-grid update): sample s is one jittered point inside grid cell cells[s] (or s)
  // of a level's R^3 grid and the result goes straight into the level's occupancy values
  const int64_t* cells;   // nullable: cell ids (x R^2 + y R + z); null = cell s
  const float* jitter;    // [n,3] in [0,1): position inside the cell
  float cell_lo[3], cell_hi[3];
  int cell_res;           // R; 0 = not in cell mode
  int cell_morton;        // cells == null: walk the level's cells in Morton order (power-of-two R up to 1024)
  int cell_update;        // 1: occs[c] = max(occs[c] * decay, occ) (every cell at most once); 2: atomic max into cand[c]
  float occ_scale, occ_decay;
  float* occs;            // the level's R^3 values (mode 1) or zero-initialised candidates (mode 2)
  uint8_t* touched;       // mode 2: cells that received a candidate
  CednerfFieldDesc d;
};

#define FIELD_MAX_GROUPS 8

// Cell mode without a cell list (the warm-up update: every cell of the level once).  Thread s does not take cell s - 32
// consecutive cell ids are a 32-cell column along z, a quarter of the grid's width, so the lanes of a warp sat in
// different cells of every hash level - but the s-th cell of a Morton walk (power-of-two resolutions): a warp is a compact
// 4 x 4 x 2 block of cells and its gathers coalesce like those of neighbouring ray samples.  The random draws stay
// attached to the CELL (jitter / timestamp number c belongs to cell c, as in nerfacc's element-wise update), so every
// cell is evaluated at exactly the same point as before.
__device__ __forceinline__ uint32_t compact3(uint32_t v) {  // every third bit of v, packed
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}
__device__ __forceinline__ int64_t cell_of(const FieldFwdArgs& a, int64_t s) {
  if (a.cells) return a.cells[s];
  if (!a.cell_morton) return s;
  const uint32_t m = (uint32_t)s;
  const int64_t R = a.cell_res;
  return ((int64_t)compact3(m >> 2) * R + compact3(m >> 1)) * R + compact3(m);  // x from bits 2, 5, ..: z fastest in the id
}
// index of sample s's random draws (jitter, timestamp): by position in the cell list, by cell id in the Morton walk
__device__ __forceinline__ int64_t draw_of(const FieldFwdArgs& a, int64_t s) { return a.cells ? s : cell_of(a, s); }

// -DFIELD_PHASE_CLOCKS (the libcednerf_b200_dbg.so of `make debug`, profiles/tools/exp_field_fwd.py): the first thread of
// every warp-group adds the cycles each phase of a tile took to g_phase_clocks; slot 7 counts tiles.
#ifdef FIELD_PHASE_CLOCKS
__device__ unsigned long long g_phase_clocks[8];
#define PHASE_MARK(i)                                                      \
  do {                                                                     \
    if (gtid == 0) {                                                       \
      const long long now_ = clock64();                                    \
      atomicAdd(&g_phase_clocks[i], (unsigned long long)(now_ - phase_t)); \
      phase_t = now_;                                                      \
    }                                                                      \
  } while (0)
#else
#define PHASE_MARK(i) do {} while (0)
#endif

// GROUPS = 128-sample tiles in flight per SM (one warp-group each); the register budget follows (7: 72, 6: 80 per thread).
// Measured on the 9.06 M-sample pre-pass (density only) 4 / 5 / 6 / 7 groups: 1.67 / 1.49 / 1.39 / 1.46 ms; on the rounds of
// a 1352 x 1014 frame (density + colour) 9.8 / 8.8 / 8.2 / 7.8 ms: six for the density-only launches, seven with colour.
#ifndef FIELD_FWD_LG
#define FIELD_FWD_LG 2   // hash levels (8 gathers each) in flight per thread
#endif
#ifndef FIELD_FWD_SIGMA_GROUPS
#define FIELD_FWD_SIGMA_GROUPS 6
#endif
template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * 128, 1) field_fwd_kernel(FieldFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const CednerfFieldDesc& d = a.d;
  const bool want_rgb = a.rgb != nullptr;
  const int n_groups = blockDim.x / MLP_TILE;           // one warp-group (128 threads) per 128-sample tile
  const int tid = threadIdx.x, warp = tid >> 5, group = tid / MLP_TILE, gtid = tid % MLP_TILE;
  uint8_t* w1 = smem;
  uint8_t* w2 = w1 + d.f1.image_bytes;
  uint8_t* w3 = w2 + d.f2.image_bytes;
  uint8_t* abuf0 = reinterpret_cast<uint8_t*>(
                       (reinterpret_cast<uintptr_t>(w3 + (want_rgb ? d.f3.image_bytes : 0)) + 1023) & ~(uintptr_t)1023) +
                   (size_t)group * MLP_TILE_BYTES;
  __shared__ uint64_t bars[FIELD_MAX_GROUPS];
  __shared__ uint32_t tmem_base_s;
  uint64_t* bar = &bars[group];

  for (int q = tid; q < d.f1.image_bytes / 16; q += blockDim.x)
    reinterpret_cast<uint4*>(w1)[q] = __ldg(reinterpret_cast<const uint4*>(a.img1) + q);
  for (int q = tid; q < d.f2.image_bytes / 16; q += blockDim.x)
    reinterpret_cast<uint4*>(w2)[q] = __ldg(reinterpret_cast<const uint4*>(a.img2) + q);
  if (want_rgb)
    for (int q = tid; q < d.f3.image_bytes / 16; q += blockDim.x)
      reinterpret_cast<uint4*>(w3)[q] = __ldg(reinterpret_cast<const uint4*>(a.img3) + q);
  const uint32_t tmem_cols = n_groups <= 1 ? 64 : (n_groups == 2 ? 128 : (n_groups <= 4 ? 256 : 512));
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  if (gtid == 0) mbar_init(bar, 1);
  if (tid == 0) fence_barrier_init();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (gtid == 0) fence_barrier_init();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s + 64u * (uint32_t)group;
  const uint32_t tmem_warp = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  int64_t n_live = a.n;
  if (a.n_dev) {
    const int64_t nd = *a.n_dev;
    n_live = nd < a.n ? nd : a.n;
  }
  const int64_t n_tiles = (n_live + MLP_TILE - 1) / MLP_TILE;
  const float ms = d.moving_step;
  const int L = d.levels.n_levels;
  const uint32_t one2 = 0x3C003C00u;  // half2(1, 1): tcnn's input padding value

  for (int64_t tile = blockIdx.x + (int64_t)gridDim.x * group; tile < n_tiles; tile += (int64_t)gridDim.x * n_groups) {
    const int64_t s = tile * MLP_TILE + gtid;
    const bool ok = s < n_live;
#ifdef FIELD_PHASE_CLOCKS
    long long phase_t = clock64();
    if (gtid == 0) atomicAdd(&g_phase_clocks[7], 1ull);
#endif
    // ---- the sample: position, time, direction (cednerf/utils.py:74-104) -------------------------------------
    float x[3] = {0.f, 0.f, 0.f}, tv = 0.f;
    if (ok) {
      if (a.ridx) {
        const int64_t r = a.ridx[s];
        const float tm = __fadd_rn(a.t0[s], a.t1[s]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          x[k] = __fadd_rn(a.rays_o[3 * r + k], __fmul_rn(__fmul_rn(a.rays_d[3 * r + k], tm), 0.5f));
        tv = a.t[r * a.t_stride];
      } else if (a.cell_res) {
        // x = aabb_lo + ((coord + jitter) / R) * (aabb_hi - aabb_lo): nerfacc's element-wise chain, op by op
        const int64_t c = cell_of(a, s);
        const int R = a.cell_res;
        const int cz = (int)(c % R), cy = (int)((c / R) % R), cx = (int)(c / ((int64_t)R * R));
        const int cc[3] = {cx, cy, cz};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float u = __fdiv_rn(__fadd_rn((float)cc[k], a.jitter[3 * draw_of(a, s) + k]), (float)R);
          x[k] = __fadd_rn(a.cell_lo[k], __fmul_rn(u, __fsub_rn(a.cell_hi[k], a.cell_lo[k])));
        }
        tv = a.t[draw_of(a, s) * a.t_stride];
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) x[k] = a.x[3 * s + k];
        tv = a.t[s * a.t_stride];
      }
    }
    // ---- deformation net input: Frequency(x, y, z, t), 4 octaves (model.py:205-213) ---------------------------
    frequency_row(abuf0, gtid, x[0], x[1], x[2], tv);
    fence_proxy_async();
    tc_fence_before();
    group_sync(group);
    PHASE_MARK(0);  // sample + Frequency row
    run_chain(d.f1, w1, abuf0, tmem_base, tmem_warp, bar, phase, gtid, group);
    PHASE_MARK(1);  // deformation net
    float xn[3], mvnorm;
    bool selector = true;
    {
      uint32_t r[16];
      tmem_ld16(tmem_warp, r);
      tmem_ld_wait();
      float mv[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        mv[k] = rnd16(r[k]) * ms;
        if (d.use_div_offsets) mv[k] = mv[k] + tanhf(rnd16(r[3 + k])) * ms;
        const float xm = x[k] + mv[k];
        xn[k] = __fdiv_rn(__fsub_rn(xm, d.aabb[k]), __fsub_rn(d.aabb[3 + k], d.aabb[k]));
        selector = selector && (xn[k] > 0.f) && (xn[k] < 1.f);
      }
      mvnorm = sqrtf(mv[0] * mv[0] + mv[1] * mv[1] + mv[2] * mv[2]);
    }
    PHASE_MARK(2);  // move, normalise, selector
    // ---- density net input: [hash 2L | time 9 | 1.0 padding] (model.py:384-403) -------------------------------
    float temb[9];
    {
      // every 32-bit word (two fp16 columns) goes straight to this thread's row of the operand tile: nothing but the
      // gathers in flight stays in registers
      auto put_word = [&](int w, uint32_t v) {
        *reinterpret_cast<uint32_t*>(abuf0 + swz(gtid, w >> 2) + ((w & 3) << 2)) = v;
      };
      const int k2 = d.f2.dim_in[0];
      // runtime loops over level groups (not unrolled: the fully unrolled 16-level body was 340 KB of SASS and the
      // kernel spent 18 % of its stall samples waiting for instructions)
      if (FIELD_FWD_LG == 4 && (L & 3) == 0) {
#pragma unroll 1
        for (int l0 = 0; l0 < L; l0 += 4) {  // 32 gathers in flight per thread; four levels = one 16-byte chunk of the row
          uint32_t f4w[4];
          hash_levels<4>(xn, a.table, d.levels, 0, f4w, l0);
          *reinterpret_cast<uint4*>(abuf0 + swz(gtid, l0 >> 2)) = make_uint4(f4w[0], f4w[1], f4w[2], f4w[3]);
        }
      } else if ((L & 1) == 0) {
#pragma unroll 1
        for (int l0 = 0; l0 < L; l0 += 2) {  // 16 gathers in flight per thread; two levels = 8 bytes of the row
          uint32_t f2w[2];
          hash_levels<2>(xn, a.table, d.levels, 0, f2w, l0);
          *reinterpret_cast<uint2*>(abuf0 + swz(gtid, l0 >> 2) + ((l0 & 2) << 2)) = make_uint2(f2w[0], f2w[1]);
        }
      } else {
#pragma unroll 1
        for (int l0 = 0; l0 < L; ++l0) {
          uint32_t f1w[1];
          hash_levels<1>(xn, a.table, d.levels, 0, f1w, l0);
          put_word(l0, f1w[0]);
        }
      }
      PHASE_MARK(3);  // hash gathers + blends
      if (d.time_mode) time_embedding(tv, mvnorm, d.time_mode, temb);  // after the gathers: keeps registers free
      int w = L;
      if (d.time_mode && d.time_before_sigma) {
        // 9 time features start at column 2L (even): pairs (e0,e1) .. (e6,e7), then (e8, 1.0)
#pragma unroll
        for (int j = 0; j < 4; ++j) put_word(L + j, pack_h2(temb[2 * j], temb[2 * j + 1]));
        put_word(L + 4, pack_h2(temb[8], 1.f));
        w = L + 5;
      }
      for (; 2 * w < k2; ++w) put_word(w, one2);  // tcnn pads the input to a multiple of 16 with 1.0
    }
    fence_proxy_async();
    tc_fence_before();
    group_sync(group);
    PHASE_MARK(4);  // time embedding, padding, fence + barrier
    run_chain(d.f2, w2, abuf0, tmem_base, tmem_warp, bar, phase, gtid, group);
    PHASE_MARK(5);  // density net
    uint32_t o2[16];
    tmem_ld16(tmem_warp, o2);
    tmem_ld_wait();
    if (ok) {
      const float sg = selector ? expf(rnd16(o2[0]) - 1.f) : 0.f;  // trunc_exp(raw - 1) * selector (model.py:414-417)
      if (a.sigma) a.sigma[s] = sg;
      if (a.cell_res) {  // occ = sigma * step;  occs = max(occs * decay, occ)   (nerfacc _update, train_real.py:324-336)
        const int64_t c = cell_of(a, s);
        const float occ = __fmul_rn(sg, a.occ_scale);
        if (a.cell_update == 1) {
          a.occs[c] = fmaxf(__fmul_rn(a.occs[c], a.occ_decay), occ);
        } else {  // duplicates possible: the largest candidate wins (non-negative floats order like their bit patterns)
          atomicMax(reinterpret_cast<int*>(a.occs + c), __float_as_int(occ));
          a.touched[c] = 1;
        }
      }
    }
    if (!want_rgb) {
      tc_fence_before();
      group_sync(group);
      PHASE_MARK(6);  // sigma out + closing barrier
      continue;
    }
    // ---- colour net input: [SH4(dir) | 15 geometry features (| time 9) | 1.0 padding] (model.py:447-466) --------
    {
      float dir[3] = {0.f, 0.f, 1.f};
      if (ok) {
        const float* dp = a.ridx ? a.rays_d + 3 * a.ridx[s] : a.dirs + 3 * s;
        dir[0] = dp[0], dir[1] = dp[1], dir[2] = dp[2];
      }
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
      for (int j = 0; j < 15; ++j) in[4 + j] = rnd16(o2[1 + j]);
      if (d.time_mode && !d.time_before_sigma) {
#pragma unroll
        for (int j = 0; j < 9; ++j) in[19 + j] = temb[j];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(abuf0 + swz(gtid, c)) =
            make_uint4(pack_h2(in[8 * c], in[8 * c + 1]), pack_h2(in[8 * c + 2], in[8 * c + 3]),
                       pack_h2(in[8 * c + 4], in[8 * c + 5]), pack_h2(in[8 * c + 6], in[8 * c + 7]));
    }
    fence_proxy_async();
    tc_fence_before();
    group_sync(group);
    run_chain(d.f3, w3, abuf0, tmem_base, tmem_warp, bar, phase, gtid, group);
    {
      uint32_t r[16];
      tmem_ld16(tmem_warp, r);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int k = 0; k < 3; ++k) a.rgb[3 * s + k] = 1.f / (1.f + expf(-rnd16(r[k])));
      }
    }
    tc_fence_before();
    group_sync(group);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, tmem_cols);
}

int check_field(const CednerfFieldDesc* d, bool want_rgb) {
  if (!d) return 0;
  const int L = d->levels.n_levels;
  if (L < 1 || L > 16) return 0;
  if (d->time_mode < 0 || d->time_mode > 2) return 0;
  const int n1 = d->f1.n_layers, n2 = d->f2.n_layers, n3 = d->f3.n_layers;
  if (n1 < 1 || n1 > MLP_MAX_LAYERS || n2 < 1 || n2 > MLP_MAX_LAYERS) return 0;
  if (d->f1.dim_in[0] != 32 || d->f1.dim_out[n1 - 1] != 16) return 0;
  const int in2 = 2 * L + ((d->time_mode && d->time_before_sigma) ? 9 : 0);
  if (d->f2.dim_in[0] != (in2 + 15) / 16 * 16 || d->f2.dim_out[n2 - 1] != 16) return 0;
  if (want_rgb) {
    if (n3 < 1 || n3 > MLP_MAX_LAYERS) return 0;
    const int in3 = 19 + ((d->time_mode && !d->time_before_sigma) ? 9 : 0);
    if (d->f3.dim_in[0] != (in3 + 15) / 16 * 16 || d->f3.dim_out[n3 - 1] != 16) return 0;
  }
  return 1;
}

}  // namespace

// sigma (and rgb) of n samples.  Samples are either packed ray samples (ray_indices, t_starts, t_ends, rays_o,
// rays_d; timestamps indexed by ray) or explicit points (x, dirs; timestamps indexed by point); t_stride 0 = one
// timestamp for all (the reference's eval path, cednerf/utils.py:187-191).  rgb == NULL: density only (the sigma_fn
// pre-pass of OccGridEstimator.sampling and occ_eval_fn).  n_device (nullable): the number of live samples is read from
// device memory (min(n, *n_device)); n is then only the capacity of the buffers - the marching rounds of
// render_image_test use it to skip the host read of each round's sample total.
CEDNERF_EXPORT int cednerf_field_fwd(const int64_t* ray_indices, const float* t_starts, const float* t_ends,
                                     const float* rays_o, const float* rays_d, const float* x, const float* dirs,
                                     const float* timestamps, int t_stride, int64_t n, const void* image_deform,
                                     const void* image_density, const void* image_colour, const void* table_f16,
                                     const CednerfFieldDesc* desc, float* sigma, float* rgb, const int64_t* n_device,
                                     void* stream) {
  CEDNERF_REQUIRE(check_field(desc, rgb != nullptr), "bad field descriptor");
  CEDNERF_REQUIRE(n >= 0 && sigma && timestamps, "bad arguments");
  CEDNERF_REQUIRE((ray_indices && t_starts && t_ends && rays_o && rays_d) || (!ray_indices && x), "need packed samples or points");
  CEDNERF_REQUIRE(!rgb || ray_indices || dirs, "colour needs directions");
  if (n == 0) return 0;
  // one CTA per SM with six (density only) or seven (with colour) 128-sample tiles in flight sharing one copy of the
  // weight images (see the kernel's header for the measurements)
  const int n_groups = rgb ? 7 : FIELD_FWD_SIGMA_GROUPS;
  const int smem = desc->f1.image_bytes + desc->f2.image_bytes + (rgb ? desc->f3.image_bytes : 0) +
                   n_groups * MLP_TILE_BYTES + 2048;
  static CednerfOncePerDevice configured7, configured6;
  if (int e = rgb ? cednerf_opt_in_smem(field_fwd_kernel<7>, 224 * 1024, configured7, "cednerf_field_fwd")
                  : cednerf_opt_in_smem(field_fwd_kernel<FIELD_FWD_SIGMA_GROUPS>, 224 * 1024, configured6, "cednerf_field_fwd"))
    return e;
  CEDNERF_REQUIRE(smem <= 224 * 1024, "networks too large for the fused kernel");
  FieldFwdArgs a{};
  a.ridx = ray_indices, a.t0 = t_starts, a.t1 = t_ends, a.rays_o = rays_o, a.rays_d = rays_d, a.x = x, a.dirs = dirs;
  a.t = timestamps, a.t_stride = t_stride, a.n = n, a.n_dev = n_device;
  a.img1 = (const uint8_t*)image_deform, a.img2 = (const uint8_t*)image_density, a.img3 = (const uint8_t*)image_colour;
  a.table = (const __half*)table_f16, a.sigma = sigma, a.rgb = rgb, a.d = *desc;
  const int64_t tiles = (n + MLP_TILE - 1) / MLP_TILE;
  const int64_t ctas = tiles;  // tiles go round-robin over CTAs first, then over the groups of a CTA
  const int64_t max_ctas = (int64_t)cednerf_num_sms();
  const unsigned grid = (unsigned)(ctas < max_ctas ? ctas : max_ctas);
  if (rgb) field_fwd_kernel<7><<<grid, 7 * MLP_TILE, smem, (cudaStream_t)stream>>>(a);
  else field_fwd_kernel<FIELD_FWD_SIGMA_GROUPS><<<grid, FIELD_FWD_SIGMA_GROUPS * MLP_TILE, smem, (cudaStream_t)stream>>>(a);
  return cednerf_check_launch("cednerf_field_fwd");
}

namespace {
__global__ void occ_finalize_kernel(float* __restrict__ occs, float* __restrict__ cand, uint8_t* __restrict__ touched,
                                    int64_t n_cells, float decay) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells || !touched[c]) return;
  occs[c] = fmaxf(__fmul_rn(occs[c], decay), cand[c]);
  cand[c] = 0.f;
  touched[c] = 0;
}
}  // namespace

// Occupancy-grid update of one level, fused (nerfacc OccGridEstimator._update as called at
// train_real.py:324-336 with occ_eval_fn(x) = query_density(x, t)["density"] * step): n jittered points, one per entry of
// `cells` (null: every cell of the level once, the warm-up case), are pushed through the deformation net, the hash
// grid and the density net, and occs[cell] = max(occs[cell] * ema_decay, sigma * step) is written by the same kernel.
// `cells` without duplicates: leave cand / touched null.  With duplicates (the uniform + occupied draws after the
// warm-up) pass zero-initialised cand [R^3] and touched [R^3]: the largest candidate of a cell wins (what scatter-amax
// gives), applied by a second small launch that also clears the two buffers again.
CEDNERF_EXPORT int cednerf_occ_update_level(const int64_t* cells, int64_t n, const float* jitter, const float* timestamps,
                                            const float* level_aabb, int resolution, float step_scale, float ema_decay,
                                            const void* image_deform, const void* image_density, const void* table_f16,
                                            const CednerfFieldDesc* desc, float* occs_level, float* cand, uint8_t* touched,
                                            void* stream) {
  CEDNERF_REQUIRE(check_field(desc, false), "bad field descriptor");
  CEDNERF_REQUIRE(n >= 0 && jitter && timestamps && level_aabb && resolution > 0 && occs_level, "bad arguments");
  CEDNERF_REQUIRE((cand == nullptr) == (touched == nullptr), "cand and touched go together");
  CEDNERF_REQUIRE(cells || n <= (int64_t)resolution * resolution * resolution, "more points than cells");
  if (n == 0) return 0;
  const int n_groups = 6;
  const int smem = desc->f1.image_bytes + desc->f2.image_bytes + n_groups * MLP_TILE_BYTES + 2048;
  static CednerfOncePerDevice configured;
  if (int e = cednerf_opt_in_smem(field_fwd_kernel<6>, 224 * 1024, configured, "cednerf_occ_update_level")) return e;
  CEDNERF_REQUIRE(smem <= 224 * 1024, "networks too large for the fused kernel");
  FieldFwdArgs a{};
  a.t = timestamps, a.t_stride = 1, a.n = n;
  a.img1 = (const uint8_t*)image_deform, a.img2 = (const uint8_t*)image_density;
  a.table = (const __half*)table_f16, a.d = *desc;
  a.cells = cells, a.jitter = jitter, a.cell_res = resolution, a.cell_update = cand ? 2 : 1;
  a.cell_morton = !cells && (resolution & (resolution - 1)) == 0 && resolution <= 1024 &&
                  n == (int64_t)resolution * resolution * resolution;
  // level_aabb is a HOST array of 6 floats (the estimator's aabbs[level])
  for (int k = 0; k < 3; ++k) a.cell_lo[k] = level_aabb[k], a.cell_hi[k] = level_aabb[3 + k];
  a.occ_scale = step_scale, a.occ_decay = ema_decay, a.occs = cand ? cand : occs_level, a.touched = touched;
  const int64_t tiles = (n + MLP_TILE - 1) / MLP_TILE;
  const int64_t max_ctas = (int64_t)cednerf_num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  field_fwd_kernel<6><<<(unsigned)(tiles < max_ctas ? tiles : max_ctas), n_groups * MLP_TILE, smem, st>>>(a);
  int launches = 1;
  if (cand) {
    const int64_t n_cells = (int64_t)resolution * resolution * resolution;
    occ_finalize_kernel<<<cednerf_blocks(n_cells, 256), 256, 0, st>>>(occs_level, cand, touched, n_cells, ema_decay);
    ++launches;
  }
  return cednerf_check_launch("cednerf_occ_update_level", launches);
}

#ifdef FIELD_PHASE_CLOCKS
// debug build only: copy the accumulated phase clocks to `out` (8 x uint64 on the host) and clear them
CEDNERF_EXPORT int cednerf_debug_phase_clocks(unsigned long long* out) {
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out, g_phase_clocks, sizeof(zero)) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(g_phase_clocks, zero, sizeof(zero)) != cudaSuccess;
}
#endif
