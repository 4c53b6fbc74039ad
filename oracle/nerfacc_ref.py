"""ORACLE (test infrastructure) — restatement of the nerfacc pieces the reference calls.

nerfacc is an un-vendored, un-pinned dependency of the reference (API shape implies
>= 0.5.3); it is absent from this container, so its published semantics are restated
from SURVEY.md Appendix A.  PARITY UNPINNED (see oracle/__init__.py).

Reference call sites:
  OccGridEstimator ctor            train_real.py:185-187
  .sampling                        cednerf/utils.py:115-125
  .update_every_n_steps            train_real.py:332-336
  traverse_grids                   cednerf/utils.py:245-264
  ray_aabb_intersect               cednerf/utils.py:215
  render_weight_from_density       cednerf/render.py:81-87, cednerf/utils.py:274-281
  render_transmittance_from_density cednerf/render.py:52-54
  accumulate_along_rays(_)         cednerf/render.py:158-169, cednerf/utils.py:282-299
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Callable, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/march_oracle.c -> oracle/libcednerf_oracle.so (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "libcednerf_oracle.so")
    src = os.path.join(_HERE, "march_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libcednerf_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f32(t):
    return t.detach().to(torch.float32).contiguous().cpu()


# --------------------------------------------------------------------------------------
# grid.py
# --------------------------------------------------------------------------------------
class RayIntervals:
    def __init__(self, vals, packed_info=None, ray_indices=None, is_left=None, is_right=None):
        self.vals, self.packed_info, self.ray_indices = vals, packed_info, ray_indices
        self.is_left, self.is_right = is_left, is_right


class RaySamples:
    def __init__(self, vals, packed_info=None, ray_indices=None, is_valid=None):
        self.vals, self.packed_info, self.ray_indices, self.is_valid = vals, packed_info, ray_indices, is_valid


def ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane=-float("inf"), far_plane=float("inf"),
                       miss_value=float("inf")):
    """Appendix A.4.  Returns t_mins[N,L], t_maxs[N,L], hits[N,L] (bool)."""
    o, d, bx = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    n, l = o.shape[0], bx.shape[0]
    t_mins = torch.empty(n, l)
    t_maxs = torch.empty(n, l)
    hits = torch.empty(n, l, dtype=torch.uint8)
    _lib().oracle_ray_aabb_intersect(_p(o), _p(d), ctypes.c_int64(n), _p(bx), ctypes.c_int(l),
                                     ctypes.c_float(near_plane), ctypes.c_float(far_plane),
                                     ctypes.c_float(miss_value), _p(t_mins), _p(t_maxs), _p(hits))
    return t_mins, t_maxs, hits.bool()


def sort_boundaries(t_mins, t_maxs):
    """Stable sort of cat([t_mins, t_maxs], -1) (cednerf/utils.py:219-225)."""
    n, l = t_mins.shape
    t_sorted = torch.empty(n, 2 * l)
    t_indices = torch.empty(n, 2 * l, dtype=torch.int64)
    _lib().oracle_sort_boundaries(_p(_f32(t_mins)), _p(_f32(t_maxs)), ctypes.c_int64(n), ctypes.c_int(l),
                                  _p(t_sorted), _p(t_indices))
    return t_sorted, t_indices


def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes=None, far_planes=None, step_size=1e-3,
                   cone_angle=0.0, traverse_steps_limit=None, over_allocate=False, rays_mask=None,
                   t_sorted=None, t_indices=None, hits=None, packed_only=False):
    """Appendix A.5/A.6.  Returns (RayIntervals, RaySamples, termination_planes).

    With ``packed_only`` the third-party interval format is skipped and
    ``(ray_indices, t_starts, t_ends, packed_info, termination_planes)`` is returned instead.
    """
    o, d, bx = _f32(rays_o), _f32(rays_d), _f32(aabbs)
    n = o.shape[0]
    nl, res = int(binaries.shape[0]), int(binaries.shape[1])
    assert binaries.shape[1] == binaries.shape[2] == binaries.shape[3], "oracle assumes cubic grids"
    bins = binaries.detach().to(torch.uint8).contiguous().cpu()
    near = torch.zeros(n) if near_planes is None else _f32(near_planes)
    far = torch.full((n,), float("inf")) if far_planes is None else _f32(far_planes)
    mask = None if rays_mask is None else rays_mask.detach().to(torch.uint8).contiguous().cpu()
    limit = -1 if traverse_steps_limit is None else int(traverse_steps_limit)
    if over_allocate:
        assert limit > 0, "over_allocate needs traverse_steps_limit > 0"
    if t_sorted is None or t_indices is None or hits is None:
        t_mins, t_maxs, hits = ray_aabb_intersect(o, d, bx)
        t_sorted, t_indices = sort_boundaries(t_mins, t_maxs)
    ts, ti = _f32(t_sorted), t_indices.detach().to(torch.int64).contiguous().cpu()
    hb = hits.detach().to(torch.uint8).contiguous().cpu()

    n_iv = torch.zeros(n, dtype=torch.int32)
    n_sm = torch.zeros(n, dtype=torch.int32)
    term = torch.zeros(n)
    L = _lib()
    common = (_p(o), _p(d), ctypes.c_int64(n), _p(bins), _p(bx), ctypes.c_int(nl), ctypes.c_int(res),
              _p(near), _p(far), ctypes.c_float(step_size), ctypes.c_float(cone_angle), ctypes.c_int(limit),
              _p(mask), _p(ts), _p(ti), _p(hb))
    if over_allocate:
        alive = torch.ones(n, dtype=torch.int64) if mask is None else mask.to(torch.int64)
        iv_cnt, sm_cnt = alive * (2 * limit), alive * limit
    else:
        L.oracle_march_count(*common, _p(n_iv), _p(n_sm), _p(term))
        iv_cnt, sm_cnt = n_iv.to(torch.int64), n_sm.to(torch.int64)
    iv_start = torch.cumsum(iv_cnt, 0) - iv_cnt
    sm_start = torch.cumsum(sm_cnt, 0) - sm_cnt
    n_iv_tot, n_sm_tot = int(iv_cnt.sum()), int(sm_cnt.sum())

    if packed_only:
        assert not over_allocate
        t0 = torch.empty(n_sm_tot)
        t1 = torch.empty(n_sm_tot)
        ridx = torch.empty(n_sm_tot, dtype=torch.int64)
        L.oracle_march_fill(*common, _p(None), _p(sm_start), _p(None), _p(None), _p(None), _p(None), _p(None),
                            _p(ridx), _p(None), _p(t0), _p(t1), _p(None), _p(None), _p(term))
        return ridx, t0, t1, torch.stack([sm_start, sm_cnt], -1), term

    iv_vals = torch.zeros(n_iv_tot)
    iv_left = torch.zeros(n_iv_tot, dtype=torch.uint8)
    iv_right = torch.zeros(n_iv_tot, dtype=torch.uint8)
    iv_ray = torch.zeros(n_iv_tot, dtype=torch.int64)
    sm_vals = torch.zeros(n_sm_tot)
    sm_ray = torch.zeros(n_sm_tot, dtype=torch.int64)
    sm_valid = torch.zeros(n_sm_tot, dtype=torch.uint8)
    L.oracle_march_fill(*common, _p(iv_start), _p(sm_start), _p(iv_vals), _p(iv_left), _p(iv_right), _p(iv_ray),
                        _p(sm_vals), _p(sm_ray), _p(sm_valid), _p(None), _p(None), _p(n_iv), _p(n_sm), _p(term))
    if over_allocate:
        # packed_info keeps the reserved chunk starts and the *actual* counts (A.5)
        iv_pack = torch.stack([iv_start, n_iv.to(torch.int64)], -1)
        sm_pack = torch.stack([sm_start, n_sm.to(torch.int64)], -1)
    else:
        iv_pack = torch.stack([iv_start, iv_cnt], -1)
        sm_pack = torch.stack([sm_start, sm_cnt], -1)
    intervals = RayIntervals(iv_vals, iv_pack, iv_ray, iv_left.bool(), iv_right.bool())
    samples = RaySamples(sm_vals, sm_pack, sm_ray, sm_valid.bool())
    return intervals, samples, term


def traverse_grids_py(rays_o, rays_d, binaries, aabbs, near_planes, far_planes, step_size, cone_angle,
                      limit=-1):
    """Independent pure-Python (numpy float32 scalar) restatement of A.6 for tiny cases.

    Used only to cross-check march_oracle.c.  Returns list of (ray, t_start, t_end) and termination planes.
    """
    f = np.float32
    o_all, d_all = _f32(rays_o).numpy(), _f32(rays_d).numpy()
    bx_all = _f32(aabbs).numpy()
    bins = binaries.detach().cpu().numpy().astype(bool)
    nl, res = bins.shape[0], bins.shape[1]
    t_mins, t_maxs, hits = ray_aabb_intersect(rays_o, rays_d, aabbs)
    t_sorted, t_indices = sort_boundaries(t_mins, t_maxs)
    t_sorted, t_indices, hits = t_sorted.numpy(), t_indices.numpy(), hits.numpy()
    step_size, cone_angle, eps = f(step_size), f(cone_angle), f(1e-6)

    def dt_of(t):
        return min(max(f(t * cone_angle), step_size), f(1e10))

    out, term = [], np.zeros(o_all.shape[0], dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for r in range(o_all.shape[0]):
            o, d = o_all[r], d_all[r]
            inv = f(1.0) / d
            near, far = f(near_planes[r]), f(far_planes[r])
            t_last, continuous, n_sm = near, False, 0
            for i in range(2 * nl - 1):
                bi = int(t_indices[r, i])
                level = bi % nl
                if not hits[r, level]:
                    continue
                if bi >= nl:
                    bn = int(t_indices[r, i + 1])
                    if bn < nl:
                        continue
                    level = bn % nl
                    if not hits[r, level]:
                        continue
                tmin, tmax = max(t_sorted[r, i], near), min(t_sorted[r, i + 1], far)
                if tmin >= tmax:
                    continue
                if not continuous:
                    if step_size <= 0:
                        t_last = tmin
                    else:
                        while True:
                            dt = dt_of(t_last)
                            if f(t_last + f(dt * f(0.5))) >= tmin:
                                break
                            t_last = f(t_last + dt)
                bmin, bmax = bx_all[level, :3], bx_all[level, 3:]
                extent = bmax - bmin
                voxel = extent / f(res)
                start = o + d * f(tmin + eps)
                end = o + d * f(tmax - eps)
                cur = np.clip(np.trunc(((start - bmin) / extent) * f(res)).astype(np.int64), 0, res - 1)
                fin = np.clip(np.trunc(((end - bmin) / extent) * f(res)).astype(np.int64), 0, res - 1)
                sidx = cur + (d > 0)
                tmax_xyz = ((bmin + ((sidx.astype(np.float32) * voxel) - start)) * inv) + tmin
                stepf = np.where(d == 0, f(0), np.where(d > 0, f(1), f(-1))).astype(np.float32)
                tdist = np.where(d == 0, tmax, tmax_xyz).astype(np.float32)
                delta = np.where(d == 0, tmax, (voxel * inv) * stepf).astype(np.float32)
                stepi = stepf.astype(np.int64)
                overflow = fin + stepi
                while limit <= 0 or n_sm < limit:
                    t_trav = min(min(tdist[0], min(tdist[1], tdist[2])), tmax)
                    if not bins[level, cur[0], cur[1], cur[2]]:
                        if step_size <= 0:
                            t_last = t_trav
                        else:
                            while True:
                                dt = dt_of(t_last)
                                if f(t_last + f(dt * f(0.5))) >= t_trav:
                                    break
                                t_last = f(t_last + dt)
                        continuous = False
                    else:
                        while limit <= 0 or n_sm < limit:
                            if step_size <= 0:
                                t_next = t_trav
                            else:
                                dt = dt_of(t_last)
                                if f(t_last + f(dt * f(0.5))) >= t_trav:
                                    break
                                t_next = f(t_last + dt)
                            out.append((r, float(t_last), float(t_next)))
                            n_sm += 1
                            continuous = True
                            t_last = t_next
                            if t_next >= t_trav:
                                break
                    ax = 0 if (tdist[0] < tdist[1] and tdist[0] < tdist[2]) else (1 if tdist[1] < tdist[2] else 2)
                    cur[ax] += stepi[ax]
                    tdist[ax] = f(tdist[ax] + delta[ax])
                    if cur[ax] == overflow[ax]:
                        break
            term[r] = t_last
    return out, torch.from_numpy(term)


# --------------------------------------------------------------------------------------
# scan.py / volrend.py  (Appendix A.7) — plain torch, autograd gives the backward
# --------------------------------------------------------------------------------------
def packed_info_from_indices(ray_indices: torch.Tensor, n_rays: int) -> torch.Tensor:
    cnt = torch.bincount(ray_indices, minlength=n_rays).to(torch.int64)
    start = torch.cumsum(cnt, 0) - cnt
    return torch.stack([start, cnt], -1)


def indices_from_packed_info(packed_info: torch.Tensor) -> torch.Tensor:
    cnt = packed_info[:, 1]
    return torch.repeat_interleave(torch.arange(packed_info.shape[0], device=cnt.device), cnt)


def exclusive_sum(inputs: torch.Tensor, packed_info=None, indices=None) -> torch.Tensor:
    """Per-ray exclusive prefix sum over flattened samples (sorted by ray).  fp64 inside."""
    if inputs.numel() == 0:
        return inputs.clone()
    if indices is None:
        indices = indices_from_packed_info(packed_info)
    x = inputs.double()
    inc = torch.cumsum(x, 0)
    exc = inc - x
    is_first = torch.ones_like(indices, dtype=torch.bool)
    is_first[1:] = indices[1:] != indices[:-1]
    first_pos = torch.nonzero(is_first).squeeze(-1)
    seg = torch.cumsum(is_first.to(torch.int64), 0) - 1
    return (exc - exc[first_pos][seg]).to(inputs.dtype)


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None,
                                      n_rays=None, prefix_trans=None):
    if ray_indices is None and packed_info is not None:
        ray_indices = indices_from_packed_info(packed_info)
    sdt = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sdt)
    trans = torch.exp(-exclusive_sum(sdt, indices=ray_indices))
    if prefix_trans is not None:
        trans = trans * prefix_trans
    return trans, alphas


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                               prefix_trans=None):
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices,
                                                      n_rays, prefix_trans)
    return trans * alphas, trans, alphas


def render_visibility_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                   early_stop_eps=1e-4, alpha_thre=0.0, prefix_trans=None):
    trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices,
                                                      n_rays, prefix_trans)
    vis = trans >= early_stop_eps
    if alpha_thre > 0:
        vis = vis & (alphas >= alpha_thre)
    return vis


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    src = weights[:, None] if values is None else weights[:, None] * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype, device=src.device)
    return out.index_add(0, ray_indices, src)


def accumulate_along_rays_(weights, values=None, ray_indices=None, outputs=None):
    src = weights[:, None] if values is None else weights[:, None] * values
    outputs.index_add_(0, ray_indices, src)


# --------------------------------------------------------------------------------------
# estimators/occ_grid.py  (Appendix A.1-A.3)
# --------------------------------------------------------------------------------------
class HostRng:
    """Random draws shared by oracle and product in parity tests (CPU generator, then moved)."""

    def __init__(self, seed: int = 42, device="cpu"):
        self.g = torch.Generator().manual_seed(seed)
        self.device = device

    def randint(self, high: int, n: int) -> torch.Tensor:
        return torch.randint(high, (n,), generator=self.g).to(self.device)

    def rand(self, *shape) -> torch.Tensor:
        return torch.rand(*shape, generator=self.g).to(self.device)


class OccGridEstimator(torch.nn.Module):
    def __init__(self, roi_aabb, resolution: int = 128, levels: int = 1):
        super().__init__()
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32)
        centre, half = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
        aabbs = torch.stack([torch.cat([centre - half * 2 ** l, centre + half * 2 ** l]) for l in range(levels)])
        self.levels, self.cells_per_lvl = levels, resolution ** 3
        self.register_buffer("resolution", torch.tensor([resolution] * 3, dtype=torch.int32))
        self.register_buffer("aabbs", aabbs)
        self.register_buffer("occs", torch.zeros(levels * self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros(levels, resolution, resolution, resolution, dtype=torch.bool))
        g = torch.arange(resolution)
        coords = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
        self.register_buffer("grid_coords", coords, persistent=False)
        self.register_buffer("grid_indices", torch.arange(self.cells_per_lvl), persistent=False)

    @torch.no_grad()
    def sampling(self, rays_o, rays_d, sigma_fn: Optional[Callable] = None, near_plane=0.0, far_plane=1e10,
                 t_min=None, t_max=None, render_step_size=1e-3, early_stop_eps=1e-4, alpha_thre=0.0,
                 stratified=False, cone_angle=0.0, jitter: Optional[torch.Tensor] = None):
        n = rays_o.shape[0]
        near = torch.full((n,), float(near_plane))
        far = torch.full((n,), float(far_plane))
        if t_min is not None:
            near = torch.clamp(near, min=t_min)
        if t_max is not None:
            far = torch.clamp(far, max=t_max)
        if stratified:
            u = torch.rand(n) if jitter is None else jitter.cpu().float()
            near = near + u * render_step_size
        ridx, t0, t1, packed, _ = traverse_grids(rays_o, rays_d, self.binaries, self.aabbs, near, far,
                                                 render_step_size, cone_angle, packed_only=True)
        if (alpha_thre > 0 or early_stop_eps > 0) and sigma_fn is not None:
            alpha_thre = min(alpha_thre, self.occs.mean().item())
            sigmas = sigma_fn(t0, t1, ridx) if t0.numel() else torch.empty(0)
            assert sigmas.shape == t0.shape
            keep = render_visibility_from_density(t0, t1, sigmas, ray_indices=ridx, n_rays=n,
                                                  early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
            ridx, t0, t1 = ridx[keep], t0[keep], t1[keep]
        return ridx, t0, t1

    @torch.no_grad()
    def mark_invisible_cells(self, K, c2w, width, height, near_plane=0.0, chunk=32 ** 3):
        """nerfacc OccGridEstimator.mark_invisible_cells, called at train_real.py:205-211 [UPSTREAM, restated from nerfacc
        0.5.x: parity unpinned].  Cells are taken at coord / (res - 1); a cell is valid (occs = 0) when some camera sees
        it at depth >= near_plane and no camera sees it closer; otherwise occs = -1.  The 3x3 products are written out
        term by term (left-to-right sums, every operation rounded in fp32) so the CUDA kernel can be bit-compared."""
        K, c2w = K.float().cpu(), c2w.float().cpu()
        n_cams = c2w.shape[0]
        rt = c2w[:, :3, :3].transpose(2, 1)                      # w2c_R
        t = c2w[:, :3, 3]
        w2c_t = -((rt[:, :, 0] * t[:, 0:1] + rt[:, :, 1] * t[:, 1:2]) + rt[:, :, 2] * t[:, 2:3])   # [C,3]
        kk = K.expand(n_cams, 3, 3)
        res = self.resolution.float()
        for lvl in range(self.levels):
            for i in range(0, self.cells_per_lvl, chunk):
                x = self.grid_coords[i:i + chunk].float() / (res - 1)
                xw = self.aabbs[lvl, :3] + x * (self.aabbs[lvl, 3:] - self.aabbs[lvl, :3])           # [M,3]
                xc = [((rt[:, j, 0:1] * xw[None, :, 0] + rt[:, j, 1:2] * xw[None, :, 1]) + rt[:, j, 2:3] * xw[None, :, 2])
                      + w2c_t[:, j:j + 1] for j in range(3)]                                            # 3 x [C,M]
                uvd = [(kk[:, j, 0:1] * xc[0] + kk[:, j, 1:2] * xc[1]) + kk[:, j, 2:3] * xc[2] for j in range(3)]
                u, v = uvd[0] / uvd[2], uvd[1] / uvd[2]
                in_image = (uvd[2] >= 0) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
                covered = ((uvd[2] >= near_plane) & in_image).any(0)
                too_near = ((uvd[2] < near_plane) & in_image).any(0)
                valid = covered & ~too_near
                base = lvl * self.cells_per_lvl + i
                self.occs[base:base + valid.numel()] = torch.where(valid, 0.0, -1.0)

    @torch.no_grad()
    def update_every_n_steps(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16,
                             rng: Optional[HostRng] = None):
        if step % n == 0 and self.training:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, rng)

    @torch.no_grad()
    def _update(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, rng=None):
        rng = rng or HostRng(step)
        cpl = self.cells_per_lvl
        lvl_indices = []
        if step < warmup_steps:
            for l in range(self.levels):
                lvl_indices.append(self.grid_indices[self.occs[l * cpl + self.grid_indices] >= 0])
        else:
            n = cpl // 4
            for l in range(self.levels):
                uni = rng.randint(cpl, n)
                uni = uni[self.occs[l * cpl + uni] >= 0]
                occ_idx = torch.nonzero(self.binaries[l].flatten())[:, 0]
                if n < len(occ_idx):
                    occ_idx = occ_idx[rng.randint(len(occ_idx), n)]
                lvl_indices.append(torch.cat([uni, occ_idx]))
        for l, idx in enumerate(lvl_indices):
            x = (self.grid_coords[idx].float() + rng.rand(len(idx), 3)) / self.resolution.float()
            x = self.aabbs[l, :3] + x * (self.aabbs[l, 3:] - self.aabbs[l, :3])
            occ = occ_eval_fn(x).squeeze(-1).float()
            cell = l * cpl + idx
            if getattr(self, "duplicate_rule", "last") == "max":
                # a cell drawn several times (uniform + occupied draws after the warm-up): nerfacc's indexed assignment
                # keeps ONE of the candidates, which one is undefined on CUDA; "max" is the deterministic choice the
                # product makes (the largest candidate), "last" what index_put does on the CPU
                uniq = torch.unique(cell)
                cand = torch.full_like(self.occs, -1.0).scatter_reduce(0, cell, occ, "amax", include_self=True)
                self.occs[uniq] = torch.maximum(self.occs[uniq] * ema_decay, cand[uniq])
                continue
            self.occs[cell] = torch.maximum(self.occs[cell] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)
