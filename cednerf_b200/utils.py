"""cednerf/utils.py surface: render_image (:46-150), render_image_test (:153-318), trunc_exp, set_random_seed."""
from __future__ import annotations

import collections
import random

import numpy as np
import torch

from . import nerfacc, ops
from .model import trunc_exp  # noqa: F401  (re-exported like cednerf/utils.py:43)
from .render import rendering

Rays = collections.namedtuple("Rays", ("origins", "viewdirs"))


def namedtuple_map(fn, tup):
    """datasets/utils.py:8-15."""
    return type(tup)(*(None if x is None else fn(x) for x in tup))


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _field_fns(field, rays, timestamps):
    def positions_of(t0, t1, ridx):
        o, d = rays.origins[ridx], rays.viewdirs[ridx]
        x = o + d * (t0 + t1)[:, None] / 2.0
        t = timestamps[ridx] if field.training else timestamps.expand_as(x[:, :1])
        return x, t, d

    def fused(t0, t1, ridx, sigma_only):
        """One launch from packed samples: positions, encodings and networks never leave the SM."""
        per_ray = field.training and timestamps.numel() == rays.origins.shape[0]
        counts = ops.counts_of(ridx)
        return field.fused_query(t0.numel(), packed=(ridx, t0, t1, rays.origins, rays.viewdirs), timestamps=timestamps,
                                 t_stride=1 if per_ray else 0, sigma_only=sigma_only,
                                 n_dev=None if counts is None else counts[1])

    def can_fuse(t0):
        ok = (not torch.is_grad_enabled()) and t0.numel() > 0 and getattr(field, "fused_supported", lambda: False)()
        return ok and (timestamps.numel() == 1 or (field.training and timestamps.numel() == rays.origins.shape[0]))

    def sigma_fn(t0, t1, ridx):
        if can_fuse(t0):
            return fused(t0, t1, ridx, True)[0]
        x, t, _ = positions_of(t0, t1, ridx)
        return field.query_density(x, t)["density"].squeeze(-1)

    def rgb_sigma_fn(t0, t1, ridx):
        if (field.training and torch.is_grad_enabled() and t0.numel() > 0
                and getattr(field, "fused_train_supported", lambda: False)()
                and (timestamps.numel() == 1 or timestamps.numel() == rays.origins.shape[0])):
            return field.fused_train(ridx, t0, t1, rays.origins, rays.viewdirs, timestamps,
                                     1 if timestamps.numel() == rays.origins.shape[0] else 0)
        if can_fuse(t0) and not field.training:
            sigma, rgb = fused(t0, t1, ridx, False)
            return rgb, {"density": sigma[:, None]}
        x, t, d = positions_of(t0, t1, ridx)
        return field(x, t, d)

    return sigma_fn, rgb_sigma_fn


def render_image(radiance_field, estimator, rays, near_plane=0.0, far_plane=1e10, render_step_size=1e-3,
                 render_bkgd=None, cone_angle=0.0, alpha_thre=0.0, test_chunk_size=8192, timestamps=None,
                 jitter=None, device_counts=False):
    """-> (rgb, acc, depth, n_rendering_samples, extras list).  `jitter` (optional, [n_rays] in [0,1)) replaces the
    stratified random draw so that parity tests do not depend on the RNG stream.

    device_counts=True (training, fused field only): nothing is read back by the host during the call.  The packed
    sample tensors in `extras` are then allocated at a capacity derived from earlier batches, only their first
    n_rendering_samples rows are live, and n_rendering_samples is a 0-dim int64 DEVICE tensor (the reference's training
    loop reads it once per step for its ray-count controller, train_real.py:354-360: `int(n)` does that)."""
    shape = rays.origins.shape
    rays = Rays(rays.origins.reshape(-1, 3), rays.viewdirs.reshape(-1, 3))
    n = rays.origins.shape[0]
    chunk = n if radiance_field.training else test_chunk_size
    outs, infos = [], []
    for i in range(0, n, chunk):
        cr = Rays(rays.origins[i:i + chunk], rays.viewdirs[i:i + chunk])
        ts = timestamps[i:i + chunk] if (radiance_field.training and timestamps is not None) else timestamps
        sigma_fn, rgb_sigma_fn = _field_fns(radiance_field, cr, ts)
        ridx, t0, t1 = estimator.sampling(cr.origins, cr.viewdirs, sigma_fn=sigma_fn, near_plane=near_plane,
                                          far_plane=far_plane, render_step_size=render_step_size,
                                          stratified=radiance_field.training, cone_angle=cone_angle,
                                          alpha_thre=alpha_thre, jitter=None if jitter is None else jitter[i:i + chunk],
                                          device_counts=bool(device_counts and radiance_field.training and
                                                             getattr(radiance_field, "fused_train_supported", lambda: False)()
                                                             and timestamps is not None and torch.is_grad_enabled()))
        rgb, opac, depth, extras = rendering(t0, t1, ridx, cr.origins.shape[0], rgb_sigma_fn=rgb_sigma_fn,
                                             render_bkgd=render_bkgd)
        extras.update(ray_indices=ridx, t_starts=t0, t_ends=t1)
        counts = ops.counts_of(ridx)
        outs.append((rgb, opac, depth, len(t0) if counts is None else counts[1].view(())))
        infos.append(extras)
    rgb, opac, depth = (torch.cat([o[k] for o in outs], 0) for k in range(3))
    return (rgb.view(*shape[:-1], -1), opac.view(*shape[:-1], -1), depth.view(*shape[:-1], -1),
            outs[0][3] if len(outs) == 1 else sum(o[3] for o in outs), infos)


@torch.no_grad()
def render_image_test(max_samples, radiance_field, estimator, rays, near_plane=0.0, far_plane=1e10,
                      render_step_size=1e-3, render_bkgd=None, cone_angle=0.0, alpha_thre=0.0, early_stop_eps=1e-4,
                      timestamps=None):
    """Iterative early-terminating marcher of cednerf/utils.py:153-318 -> (rgb, acc, depth, n_samples)."""
    shape = rays.origins.shape
    rays = Rays(rays.origins.reshape(-1, 3), rays.viewdirs.reshape(-1, 3))
    n, dev = rays.origins.shape[0], rays.origins.device
    _, rgb_sigma_fn = _field_fns(radiance_field, rays, timestamps)
    opacity, depth, rgb = torch.zeros(n, 1, device=dev), torch.zeros(n, 1, device=dev), torch.zeros(n, 3, device=dev)
    alive = torch.ones(n, dtype=torch.bool, device=dev)
    min_samples = 1 if cone_angle == 0 else 4
    near = torch.full((n,), float(near_plane), device=dev)
    far = torch.full((n,), float(far_plane), device=dev)
    t_mins, t_maxs, hits = nerfacc.ray_aabb_intersect(rays.origins, rays.viewdirs, estimator.aabbs)
    t_sorted, t_indices = ops.sort_boundaries(t_mins, t_maxs)
    bits = nerfacc.grid.occupancy_bits(estimator.binaries)
    res = int(estimator.binaries.shape[1])
    done = total = 0
    # With the fused field kernel a round needs ONE host read (the reference's `alive.sum()`): the round's sample total
    # stays on the device - outputs are sized for the bound n_alive * k and the kernels read the live count themselves.
    fuse = (getattr(radiance_field, "fused_supported", lambda: False)() and not radiance_field.training
            and timestamps is not None and timestamps.numel() == 1 and render_step_size > 0)
    total_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    while done < max_samples:
        n_alive = int(alive.sum())                     # the reference's host read (utils.py:231)
        if n_alive == 0:
            break
        k = max(min(n // n_alive, 64), min_samples)
        done += k
        # traverse_grids(..., k, over_allocate=True, alive, ...) of utils.py:245-264, emitted directly in packed form:
        # same samples, no interval flags to compact afterwards
        mi = ops.MarchInputs(rays.origins, rays.viewdirs, bits, estimator.aabbs, res, near, far, 0.0, float("inf"),
                             render_step_size, cone_angle, k, alive, t_sorted, t_indices, hits)
        _, n_sm, term = mi.count(record_runs=True)
        starts, _, n_tot = ops.exclusive_scan(n_sm, want_packed=False)
        if fuse and mi.runs is not None:
            cap = n_alive * k
            ridx, t0, t1 = mi.fill_packed_from_runs(starts, cap)
            sigma, rgbs = radiance_field.fused_query(cap, packed=(ridx, t0, t1, rays.origins, rays.viewdirs),
                                                     timestamps=timestamps, t_stride=0, sigma_only=False, n_dev=n_tot)
            ops.composite_round_(t0, t1, sigma, rgbs, torch.cat([starts, n_tot]), rgb, opacity, depth)
            total_dev += n_tot
        else:
            n_round = int(n_tot.item())
            if n_round:
                if mi.runs is not None:
                    ridx, t0, t1 = mi.fill_packed_from_runs(starts, n_round)
                else:
                    ridx, t0, t1, _ = mi.fill_packed(starts, n_round)
                rgbs, sres = rgb_sigma_fn(t0, t1, ridx)
                ops.composite_round_(t0, t1, sres["density"].squeeze(-1), rgbs, torch.cat([starts, n_tot]), rgb, opacity,
                                     depth)
            total += n_round
        near = term
        alive = (opacity.view(-1) <= 1 - early_stop_eps) & (n_sm == k)
    total += int(total_dev.item())
    rgb = rgb + render_bkgd * (1.0 - opacity)
    depth = depth / opacity.clamp_min(torch.finfo(torch.float32).eps)
    return rgb.view(*shape[:-1], -1), opacity.view(*shape[:-1], -1), depth.view(*shape[:-1], -1), total
