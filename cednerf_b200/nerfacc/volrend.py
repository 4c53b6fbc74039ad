"""nerfacc.volrend surface (reference call sites cednerf/render.py:52-54, :81-87, :158-169;
cednerf/utils.py:274-299)."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops


def _offsets(n_samples, packed_info, ray_indices, n_rays):
    if ray_indices is not None:
        if n_rays is None:
            raise ValueError("n_rays is required with ray_indices")
        return ops.ray_offsets(ray_indices, int(n_rays)), int(n_rays)
    if packed_info is None:  # a single ray
        dev = torch.cuda.current_device()
        return torch.tensor([0, n_samples], dtype=torch.int64, device=f"cuda:{dev}"), 1
    cnt = packed_info[:, 1].to(torch.int64)
    off = torch.zeros(cnt.numel() + 1, dtype=torch.int64, device=cnt.device)
    torch.cumsum(cnt, 0, out=off[1:])
    return off, cnt.numel()


def render_weight_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                               prefix_trans: Optional[torch.Tensor] = None):
    """-> (weights, trans, alphas)."""
    off, n = _offsets(t_starts.numel(), packed_info, ray_indices, n_rays)
    return ops.RenderWeightFunction.apply(t_starts, t_ends, sigmas, off, n, prefix_trans)


def render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                      prefix_trans: Optional[torch.Tensor] = None):
    """-> (trans, alphas)."""
    _, trans, alphas = render_weight_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays,
                                                  prefix_trans)
    return trans, alphas


@torch.no_grad()
def render_visibility_from_density(t_starts, t_ends, sigmas, packed_info=None, ray_indices=None, n_rays=None,
                                   early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, prefix_trans=None):
    if prefix_trans is not None:
        trans, alphas = render_transmittance_from_density(t_starts, t_ends, sigmas, packed_info, ray_indices, n_rays,
                                                          prefix_trans)
        vis = trans >= early_stop_eps
        return vis & (alphas >= alpha_thre) if alpha_thre > 0 else vis
    off, n = _offsets(t_starts.numel(), packed_info, ray_indices, n_rays)
    return ops.visibility_mask(t_starts, t_ends, sigmas, off, n, early_stop_eps, alpha_thre)


def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
    """-> [n_rays, C] (C = 1 when values is None)."""
    if ray_indices is None:
        raise NotImplementedError("accumulate_along_rays needs ray_indices (the reference always passes them)")
    ridx = ray_indices.detach().to(torch.int64).contiguous()
    off = ops.ray_offsets(ridx, int(n_rays))
    return ops.AccumulateFunction.apply(weights, values, ridx, off, int(n_rays))


@torch.no_grad()
def accumulate_along_rays_(weights, values=None, ray_indices=None, outputs=None):
    """In-place variant (cednerf/utils.py:282-299)."""
    off = ops.ray_offsets(ray_indices, outputs.shape[0])
    ops.accumulate_inplace(weights, values, off, outputs)
