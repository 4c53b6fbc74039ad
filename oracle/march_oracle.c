/*
 * ORACLE — test infrastructure only.  Never imported by the product path.
 *
 * Plain-C, single-precision restatement of the occupancy-grid ray marcher the
 * reference calls through nerfacc (un-vendored dependency, no pinned version;
 * API shape implies nerfacc >= 0.5.3):
 *
 *   - ray_aabb_intersect          reference call site cednerf/utils.py:215
 *   - sort of the 2L boundaries   reference call site cednerf/utils.py:219-225
 *   - traverse_grids              reference call sites cednerf/utils.py:241-264 (eval,
 *                                 steps-limit + over-allocate + mask) and, through
 *                                 OccGridEstimator.sampling, cednerf/utils.py:115-125 (train)
 *
 * The algorithm is restated from SURVEY.md Appendix A.4-A.6 (published nerfacc
 * semantics).  PARITY UNPINNED: the reference ships no tests or golden vectors
 * for this path and nerfacc is not installable here, so this file is the pin.
 *
 * Floating-point contract (shared with cednerf_b200/csrc/march.cu): every
 * operation is an individually rounded IEEE fp32 op in the order written here;
 * compile with -ffp-contract=off (the CUDA side uses __fmul_rn/__fadd_rn...).
 *
 * Build: see oracle/Makefile  ->  oracle/libcednerf_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORACLE_MAX_LEVELS 8

static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* A.4: slab test of one ray against L boxes.  aabbs: [L,6] = (min xyz, max xyz). */
void oracle_ray_aabb_intersect(const float* rays_o, const float* rays_d, int64_t n_rays,
                               const float* aabbs, int n_levels, float near_plane, float far_plane,
                               float miss_value, float* t_mins, float* t_maxs, uint8_t* hits) {
  for (int64_t r = 0; r < n_rays; ++r) {
    const float* o = rays_o + 3 * r;
    const float* d = rays_d + 3 * r;
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    for (int l = 0; l < n_levels; ++l) {
      const float* bx = aabbs + 6 * l;
      float tmin = -INFINITY, tmax = INFINITY;
      for (int a = 0; a < 3; ++a) {
        float t1 = (bx[a] - o[a]) * inv[a];
        float t2 = (bx[3 + a] - o[a]) * inv[a];
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
      }
      int hit = (tmax > tmin) && (tmax > 0.0f);
      hits[r * n_levels + l] = (uint8_t)hit;
      if (hit) {
        t_mins[r * n_levels + l] = clampf(tmin, near_plane, far_plane);
        t_maxs[r * n_levels + l] = clampf(tmax, near_plane, far_plane);
      } else {
        t_mins[r * n_levels + l] = miss_value;
        t_maxs[r * n_levels + l] = miss_value;
      }
    }
  }
}

/* Stable ascending sort of cat([t_mins, t_maxs]) per ray (ties keep the lower original slot). */
void oracle_sort_boundaries(const float* t_mins, const float* t_maxs, int64_t n_rays, int n_levels,
                            float* t_sorted, int64_t* t_indices) {
  const int m = 2 * n_levels;
  for (int64_t r = 0; r < n_rays; ++r) {
    float v[2 * ORACLE_MAX_LEVELS];
    int id[2 * ORACLE_MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
      v[l] = t_mins[r * n_levels + l];
      v[n_levels + l] = t_maxs[r * n_levels + l];
    }
    for (int i = 0; i < m; ++i) id[i] = i;
    for (int i = 1; i < m; ++i) { /* insertion sort = stable */
      float kv = v[i];
      int ki = id[i];
      int j = i - 1;
      while (j >= 0 && v[j] > kv) {
        v[j + 1] = v[j];
        id[j + 1] = id[j];
        --j;
      }
      v[j + 1] = kv;
      id[j + 1] = ki;
    }
    for (int i = 0; i < m; ++i) {
      t_sorted[r * m + i] = v[i];
      t_indices[r * m + i] = id[i];
    }
  }
}

static inline float step_dt(float t, float cone_angle, float step_size) {
  return clampf(t * cone_angle, step_size, 1e10f);
}

typedef struct {
  /* nerfacc-shaped interval outputs (may be NULL) */
  float* iv_vals;
  uint8_t* iv_left;
  uint8_t* iv_right;
  int64_t* iv_ray;
  /* nerfacc-shaped sample outputs (may be NULL) */
  float* sm_vals;
  int64_t* sm_ray;
  uint8_t* sm_valid;
  /* packed outputs (may be NULL) */
  float* t_starts;
  float* t_ends;
} march_out_t;

/*
 * A.6: marches one ray.  When `out` is NULL only counts are produced.
 * iv_base / sm_base are this ray's first slots in the interval / sample arrays.
 */
static void march_one_ray(int64_t r, const float* rays_o, const float* rays_d, const uint8_t* binaries,
                          const float* aabbs, int n_levels, int res, float near, float far,
                          float step_size, float cone_angle, int limit, const float* t_sorted,
                          const int64_t* t_indices, const uint8_t* hits, const march_out_t* out,
                          int64_t iv_base, int64_t sm_base, int32_t* n_iv_out, int32_t* n_sm_out,
                          float* t_term_out) {
  const float eps = 1e-6f;
  const int m = 2 * n_levels;
  const float o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
  const float d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
  const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
  const float fres = (float)res;
  const int64_t cells_per_level = (int64_t)res * res * res;

  int n_iv = 0, n_sm = 0;
  float t_last = near;
  int continuous = 0;

  for (int i = 0; i < m - 1; ++i) {
    int64_t bi = t_indices[r * m + i];
    int entering = bi < n_levels;
    int level = (int)(bi % n_levels);
    if (!hits[r * n_levels + level]) continue;
    if (!entering) {
      int64_t bn = t_indices[r * m + i + 1];
      if (bn < n_levels) continue; /* gap between boxes */
      level = (int)(bn % n_levels);
      if (!hits[r * n_levels + level]) continue;
    }
    float this_tmin = fmaxf(t_sorted[r * m + i], near);
    float this_tmax = fminf(t_sorted[r * m + i + 1], far);
    if (this_tmin >= this_tmax) continue;

    if (!continuous) {
      if (step_size <= 0.0f) {
        t_last = this_tmin;
      } else {
        for (;;) {
          float dt = step_dt(t_last, cone_angle, step_size);
          if (t_last + dt * 0.5f >= this_tmin) break;
          t_last += dt;
        }
      }
    }

    /* DDA set-up inside grid `level` */
    const float* bx = aabbs + 6 * level;
    float voxel[3], tdist[3], delta[3];
    int cur[3], overflow[3], stepi[3];
    const float ts = this_tmin + eps, te = this_tmax - eps;
    for (int a = 0; a < 3; ++a) {
      float extent = bx[3 + a] - bx[a];
      voxel[a] = extent / fres;
      float start = o[a] + d[a] * ts;
      float end = o[a] + d[a] * te;
      int c = (int)(((start - bx[a]) / extent) * fres);
      int f = (int)(((end - bx[a]) / extent) * fres);
      c = clampi(c, 0, res - 1);
      f = clampi(f, 0, res - 1);
      int start_idx = c + (d[a] > 0.0f ? 1 : 0);
      float tmax_a = ((bx[a] + (((float)start_idx * voxel[a]) - start)) * inv[a]) + this_tmin;
      float stepf = (d[a] == 0.0f) ? 0.0f : (d[a] > 0.0f ? 1.0f : -1.0f);
      tdist[a] = (d[a] == 0.0f) ? this_tmax : tmax_a;
      delta[a] = (d[a] == 0.0f) ? this_tmax : (voxel[a] * inv[a]) * stepf;
      stepi[a] = (int)stepf;
      cur[a] = c;
      overflow[a] = f + stepi[a];
    }

    while (limit <= 0 || n_sm < limit) {
      float t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
      int64_t cell = (int64_t)cur[0] * res * res + (int64_t)cur[1] * res + cur[2] + level * cells_per_level;
      if (!binaries[cell]) {
        if (step_size <= 0.0f) {
          t_last = t_trav;
        } else {
          for (;;) {
            float dt = step_dt(t_last, cone_angle, step_size);
            if (t_last + dt * 0.5f >= t_trav) break;
            t_last += dt;
          }
        }
        continuous = 0;
      } else {
        while (limit <= 0 || n_sm < limit) {
          float t_next;
          if (step_size <= 0.0f) {
            t_next = t_trav;
          } else {
            float dt = step_dt(t_last, cone_angle, step_size);
            if (t_last + dt * 0.5f >= t_trav) break;
            t_next = t_last + dt;
          }
          if (out) {
            if (out->iv_vals) {
              if (!continuous) {
                int64_t k = iv_base + n_iv;
                out->iv_vals[k] = t_last;
                out->iv_ray[k] = r;
                out->iv_left[k] = 1;
                out->iv_vals[k + 1] = t_next;
                out->iv_ray[k + 1] = r;
                out->iv_right[k + 1] = 1;
              } else {
                int64_t k = iv_base + n_iv;
                out->iv_vals[k] = t_next;
                out->iv_ray[k] = r;
                out->iv_left[k - 1] = 1;
                out->iv_right[k] = 1;
              }
            }
            if (out->sm_vals) {
              out->sm_vals[sm_base + n_sm] = (t_next + t_last) * 0.5f;
              out->sm_ray[sm_base + n_sm] = r;
              out->sm_valid[sm_base + n_sm] = 1;
            }
            if (out->t_starts) {
              out->t_starts[sm_base + n_sm] = t_last;
              out->t_ends[sm_base + n_sm] = t_next;
              if (!out->sm_vals && out->sm_ray) out->sm_ray[sm_base + n_sm] = r;
            }
          }
          n_iv += continuous ? 1 : 2;
          n_sm += 1;
          continuous = 1;
          t_last = t_next;
          if (t_next >= t_trav) break;
        }
      }
      /* advance one voxel along the axis with the nearest boundary */
      int ax = (tdist[0] < tdist[1] && tdist[0] < tdist[2]) ? 0 : (tdist[1] < tdist[2] ? 1 : 2);
      cur[ax] += stepi[ax];
      tdist[ax] += delta[ax];
      if (cur[ax] == overflow[ax]) break;
    }
  }
  *n_iv_out = n_iv;
  *n_sm_out = n_sm;
  if (t_term_out) *t_term_out = t_last;
}

/* Pass 1: per-ray interval and sample counts (+ termination planes). */
void oracle_march_count(const float* rays_o, const float* rays_d, int64_t n_rays, const uint8_t* binaries,
                        const float* aabbs, int n_levels, int res, const float* near, const float* far,
                        float step_size, float cone_angle, int limit, const uint8_t* mask,
                        const float* t_sorted, const int64_t* t_indices, const uint8_t* hits,
                        int32_t* n_intervals, int32_t* n_samples, float* termination) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n_rays; ++r) {
    if (mask && !mask[r]) {
      n_intervals[r] = 0;
      n_samples[r] = 0;
      /* nerfacc leaves termination_planes uninitialised for masked rays; the oracle keeps near */
      if (termination) termination[r] = near[r];
      continue;
    }
    march_one_ray(r, rays_o, rays_d, binaries, aabbs, n_levels, res, near[r], far[r], step_size, cone_angle,
                  limit, t_sorted, t_indices, hits, NULL, 0, 0, &n_intervals[r], &n_samples[r],
                  termination ? &termination[r] : NULL);
  }
}

/* Pass 2: fill.  iv_starts / sm_starts are the exclusive prefix sums chosen by the caller. */
void oracle_march_fill(const float* rays_o, const float* rays_d, int64_t n_rays, const uint8_t* binaries,
                       const float* aabbs, int n_levels, int res, const float* near, const float* far,
                       float step_size, float cone_angle, int limit, const uint8_t* mask,
                       const float* t_sorted, const int64_t* t_indices, const uint8_t* hits,
                       const int64_t* iv_starts, const int64_t* sm_starts, float* iv_vals, uint8_t* iv_left,
                       uint8_t* iv_right, int64_t* iv_ray, float* sm_vals, int64_t* sm_ray, uint8_t* sm_valid,
                       float* t_starts, float* t_ends, int32_t* n_intervals, int32_t* n_samples,
                       float* termination) {
  march_out_t out = {iv_vals, iv_left, iv_right, iv_ray, sm_vals, sm_ray, sm_valid, t_starts, t_ends};
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n_rays; ++r) {
    int32_t ni = 0, ns = 0;
    if (mask && !mask[r]) {
      if (n_intervals) n_intervals[r] = 0;
      if (n_samples) n_samples[r] = 0;
      if (termination) termination[r] = near[r];
      continue;
    }
    march_one_ray(r, rays_o, rays_d, binaries, aabbs, n_levels, res, near[r], far[r], step_size, cone_angle,
                  limit, t_sorted, t_indices, hits, &out, iv_starts ? iv_starts[r] : 0,
                  sm_starts ? sm_starts[r] : 0, &ni, &ns, termination ? &termination[r] : NULL);
    if (n_intervals) n_intervals[r] = ni;
    if (n_samples) n_samples[r] = ns;
  }
}
