"""nerfacc.grid surface: ray_aabb_intersect, traverse_grids (reference call sites cednerf/utils.py:215, :245-264)."""
from __future__ import annotations

import math
import weakref
from typing import Optional

import torch

from .. import ops


class RayIntervals:
    def __init__(self, vals, packed_info=None, ray_indices=None, is_left=None, is_right=None):
        self.vals, self.packed_info, self.ray_indices = vals, packed_info, ray_indices
        self.is_left, self.is_right = is_left, is_right


class RaySamples:
    def __init__(self, vals, packed_info=None, ray_indices=None, is_valid=None):
        self.vals, self.packed_info, self.ray_indices, self.is_valid = vals, packed_info, ray_indices, is_valid


def ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane: float = -math.inf, far_plane: float = math.inf,
                       miss_value: float = math.inf):
    """-> t_mins [N,L], t_maxs [N,L], hits bool [N,L]."""
    return ops.ray_aabb_intersect(rays_o, rays_d, aabbs, near_plane, far_plane, miss_value)


# Bit-packed copies of `binaries` tensors, keyed by the tensor OBJECT (id + weak reference) and validated by
# (storage, version, shape): a second estimator whose buffer happens to land on the same allocator block with the same
# version counter can never hit another tensor's entry, and entries die with their tensor.
_BITS_CACHE = {}


def _bits_key(binaries: torch.Tensor):
    return (binaries.data_ptr(), binaries._version, tuple(binaries.shape))


def _remember(binaries: torch.Tensor, bits: torch.Tensor) -> None:
    ident = id(binaries)
    _BITS_CACHE[ident] = (weakref.ref(binaries, lambda _r, i=ident: _BITS_CACHE.pop(i, None)), _bits_key(binaries), bits)


def occupancy_bits(binaries: torch.Tensor) -> torch.Tensor:
    """Bit-packed copy of `binaries`, cached per tensor object and version so repeated marches do not re-pack."""
    hit = _BITS_CACHE.get(id(binaries))
    if hit is not None and hit[0]() is binaries and hit[1] == _bits_key(binaries):
        return hit[2]
    bits = ops.pack_occupancy(binaries)
    _remember(binaries, bits)
    return bits


def set_occupancy_bits(binaries: torch.Tensor, bits: torch.Tensor) -> None:
    """Install a bit field produced together with `binaries` by a raw kernel write: the write went through a raw
    pointer, so the tensor's version counter is bumped here (anything else keyed on it must see the change)."""
    torch.autograd.graph.increment_version(binaries)
    _remember(binaries, bits)


@torch.no_grad()
def traverse_grids(rays_o, rays_d, binaries, aabbs, near_planes: Optional[torch.Tensor] = None,
                   far_planes: Optional[torch.Tensor] = None, step_size: float = 1e-3, cone_angle: float = 0.0,
                   traverse_steps_limit: Optional[int] = None, over_allocate: bool = False,
                   rays_mask: Optional[torch.Tensor] = None, t_sorted=None, t_indices=None, hits=None):
    """-> (RayIntervals, RaySamples, termination_planes); semantics of SURVEY.md Appendix A.5/A.6."""
    limit = -1 if traverse_steps_limit is None else int(traverse_steps_limit)
    if over_allocate and limit <= 0:
        raise ValueError("over_allocate needs traverse_steps_limit > 0")
    if binaries.shape[1] != binaries.shape[2] or binaries.shape[2] != binaries.shape[3]:
        raise NotImplementedError("cubic occupancy grids only")
    mi = ops.MarchInputs(rays_o, rays_d, occupancy_bits(binaries), aabbs, binaries.shape[1], near_planes, far_planes,
                         0.0, math.inf, step_size, cone_angle, limit, rays_mask, t_sorted, t_indices, hits)
    n, dev = mi.n, mi.o.device
    if over_allocate:
        alive = torch.ones(n, dtype=torch.int64, device=dev) if mi.mask is None else mi.mask.to(torch.int64)
        iv_cnt, sm_cnt = alive * (2 * limit), alive * limit
        iv_start, sm_start = torch.cumsum(iv_cnt, 0) - iv_cnt, torch.cumsum(sm_cnt, 0) - sm_cnt
        n_alive = int(alive.sum())
        n_iv_tot, n_sm_tot = n_alive * 2 * limit, n_alive * limit
    else:
        n_iv, n_sm, _ = mi.count()
        iv_start, iv_pack, iv_total = ops.exclusive_scan(n_iv)
        sm_start, sm_pack, sm_total = ops.exclusive_scan(n_sm)
        n_iv_tot, n_sm_tot = (int(v) for v in torch.cat([iv_total, sm_total]).tolist())
    iv, sm, n_iv, n_sm, term = mi.fill_nerfacc(iv_start, sm_start, n_iv_tot, n_sm_tot)
    if over_allocate:
        iv_pack = torch.stack([iv_start, n_iv.to(torch.int64)], -1)
        sm_pack = torch.stack([sm_start, n_sm.to(torch.int64)], -1)
    intervals = RayIntervals(iv[0], iv_pack, iv[3], iv[1], iv[2])
    samples = RaySamples(sm[0], sm_pack, sm[1], sm[2])
    return intervals, samples, term
