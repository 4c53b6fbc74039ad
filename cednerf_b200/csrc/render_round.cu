// One marching round of render_image_test (cednerf/utils.py:224-304) behind ONE call: count pass over the alive list ->
// capped scan -> fill from the recorded runs (+ the full-march fill of the rays over the run limit) -> fused field kernel
// with the device-side total -> compositing with the alive flag per slot -> scan + ordered compaction into the next
// list (which also begins the next round).  The eight entry points stay exported; this one exists because a frame is
// 40-220 rounds and the host paid eight ctypes calls with ~25 arguments each per round - with the kernels of this round
// at 0.3-0.5 ms the rounds of interleaved frames had become bound by the host (15 ms per 1352 x 1014 frame at 49 us per
// call; the 800 x 800 D-NeRF-shaped frame, 1380 calls, between 33 and 94 ms depending on the box's CPU).
#include "field_common.cuh"

struct CednerfRenderRound {
  const float* rays_o;
  const float* rays_d;
  int64_t n_rays;
  const uint32_t* occ_bits;
  const float* aabbs;
  int n_levels, resolution;
  float* near_term;
  float far_const, step_size, cone_angle, early_stop_eps;
  const float* t_sorted;
  const int64_t* t_indices;
  const uint8_t* hits;
  int32_t* state;
  int64_t* total;
  int32_t* n_samples;
  float* run_t;
  int32_t* run_n;
  int32_t* n_runs;
  const uint32_t* occ_coarse;
  int64_t capacity;
  int64_t* offsets;
  int64_t* totals;
  void* scan_workspace;
  float* t_starts;
  float* t_ends;
  int64_t* ray_indices;
  uint8_t* overflow;
  const float* timestamps;
  const void* image_deform;
  const void* image_density;
  const void* image_colour;
  const void* table_f16;
  const CednerfFieldDesc* desc;
  float* sigma;
  float* rgbs;
  float* colors;
  float* opacity;
  float* depth;
  int32_t* alive_flags;
  int64_t* positions;
  int64_t* position_totals;
  int run_cap, k_hint, max_samples, min_samples;
};

extern "C" {
int cednerf_march_round(int fill, const float* rays_o, const float* rays_d, int64_t n_bound, const uint32_t* occ_bits,
                        const float* aabbs, int n_levels, int resolution, float* near_term, float far_const, float step_size,
                        float cone_angle, const float* t_sorted, const int64_t* t_indices, const uint8_t* hits,
                        const int32_t* alive, const int32_t* state, const uint8_t* slot_mask, const int64_t* offsets,
                        float* t_starts, float* t_ends, int64_t* ray_indices, int32_t* n_samples, float* run_t, int32_t* run_n,
                        int32_t* n_runs, int run_cap, const uint32_t* occ_coarse, void* stream);
int cednerf_exclusive_scan_capped(const int32_t* counts, int64_t n, int64_t capacity, int64_t* offsets, int64_t* totals,
                                  void* workspace, void* stream);
int cednerf_march_fill_runs_round(int64_t n_bound, const int64_t* offsets, const int32_t* n_samples, const float* run_t,
                                  const int32_t* run_n, const int32_t* n_runs, int run_cap, float step_size, float cone_angle,
                                  const int32_t* alive, const int32_t* state, float* t_starts, float* t_ends,
                                  int64_t* ray_indices, uint8_t* overflow, void* stream);
int cednerf_field_fwd(const int64_t* ray_indices, const float* t_starts, const float* t_ends, const float* rays_o,
                      const float* rays_d, const float* x, const float* dirs, const float* timestamps, int t_stride, int64_t n,
                      const void* image_deform, const void* image_density, const void* image_colour, const void* table_f16,
                      const CednerfFieldDesc* desc, float* sigma, float* rgb, const int64_t* n_device, void* stream);
int cednerf_render_round_composite(const float* t_starts, const float* t_ends, const float* sigmas, const float* rgbs,
                                   const int64_t* offsets, const int32_t* alive, int32_t* round_state,
                                   const int32_t* slot_counts, int64_t n_bound, int k_hint, float early_stop_eps, float* colors,
                                   float* opacity, float* depth, int32_t* alive_flags, void* stream);
int cednerf_render_round_compact(const int32_t* alive_flags, const int64_t* positions, const int32_t* alive, int64_t n_bound,
                                 int32_t* round_state, int32_t* next_alive, int64_t n_rays, int max_samples, int min_samples,
                                 const int64_t* round_totals, int64_t* total, void* stream);
}

CEDNERF_EXPORT int64_t cednerf_render_round_bytes(void) { return (int64_t)sizeof(CednerfRenderRound); }

// `alive` / `next_alive`: this round's list of alive rays and the one the round writes; n_bound: an upper bound of the
// live part of `alive` (the state block holds the exact count).  Everything else is the same for every round of a frame.
CEDNERF_EXPORT int cednerf_render_round(const CednerfRenderRound* r, int64_t n_bound, const int32_t* alive,
                                        int32_t* next_alive, void* stream) {
  CEDNERF_REQUIRE(r && alive && next_alive && n_bound >= 0, "bad arguments");
  int rc;
  if ((rc = cednerf_march_round(0, r->rays_o, r->rays_d, n_bound, r->occ_bits, r->aabbs, r->n_levels, r->resolution,
                                r->near_term, r->far_const, r->step_size, r->cone_angle, r->t_sorted, r->t_indices, r->hits,
                                alive, r->state, nullptr, nullptr, nullptr, nullptr, nullptr, r->n_samples, r->run_t, r->run_n,
                                r->n_runs, r->run_cap, r->occ_coarse, stream)))
    return rc;
  if ((rc = cednerf_exclusive_scan_capped(r->n_samples, n_bound, r->capacity, r->offsets, r->totals, r->scan_workspace, stream)))
    return rc;
  if ((rc = cednerf_march_fill_runs_round(n_bound, r->offsets, r->n_samples, r->run_t, r->run_n, r->n_runs, r->run_cap,
                                          r->step_size, r->cone_angle, alive, r->state, r->t_starts, r->t_ends, r->ray_indices,
                                          r->overflow, stream)))
    return rc;
  if ((rc = cednerf_march_round(1, r->rays_o, r->rays_d, n_bound, r->occ_bits, r->aabbs, r->n_levels, r->resolution,
                                r->near_term, r->far_const, r->step_size, r->cone_angle, r->t_sorted, r->t_indices, r->hits,
                                alive, r->state, r->overflow, r->offsets, r->t_starts, r->t_ends, r->ray_indices, nullptr,
                                nullptr, nullptr, nullptr, r->run_cap, r->occ_coarse, stream)))
    return rc;
  if ((rc = cednerf_field_fwd(r->ray_indices, r->t_starts, r->t_ends, r->rays_o, r->rays_d, nullptr, nullptr, r->timestamps, 0,
                              r->capacity, r->image_deform, r->image_density, r->image_colour, r->table_f16, r->desc, r->sigma,
                              r->rgbs, r->totals, stream)))
    return rc;
  if ((rc = cednerf_render_round_composite(r->t_starts, r->t_ends, r->sigma, r->rgbs, r->offsets, alive, r->state, r->n_samples,
                                           n_bound, r->k_hint, r->early_stop_eps, r->colors, r->opacity, r->depth,
                                           r->alive_flags, stream)))
    return rc;
  if ((rc = cednerf_exclusive_scan_capped(r->alive_flags, n_bound, r->n_rays, r->positions, r->position_totals,
                                          r->scan_workspace, stream)))
    return rc;
  return cednerf_render_round_compact(r->alive_flags, r->positions, alive, n_bound, r->state, next_alive, r->n_rays,
                                      r->max_samples, r->min_samples, r->totals, r->total, stream);
}
