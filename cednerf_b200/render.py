"""cednerf/render.py surface: rendering() (cednerf/render.py:58-176) and reduce_along_rays (:8-39)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def reduce_along_rays(ray_indices, values, n_rays, weights=None, reduce: str = "mean"):
    """scatter_reduce_ with include_self=True, exactly as cednerf/render.py:8-39 ('mean' divides by count+1)."""
    src = values if weights is None else weights * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype, device=src.device)
    if ray_indices.numel() == 0:
        return out
    index = ray_indices[:, None].long().expand(-1, src.shape[-1])
    return out.scatter_reduce(0, index, src, reduce=reduce)


def rendering(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn=None, rgb_alpha_fn=None, render_bkgd=None):
    """-> (colors [n,3], opacities [n,1], depths [n,1], extras)."""
    if rgb_sigma_fn is None or rgb_alpha_fn is not None:
        raise NotImplementedError("the reference calls rendering() with rgb_sigma_fn only")
    rgbs, sres = rgb_sigma_fn(t_starts, t_ends, ray_indices)
    sigmas = sres["density"]
    assert rgbs.shape[-1] == 3, f"rgbs must have 3 channels, got {tuple(rgbs.shape)}"
    sigmas = sigmas.squeeze(-1)
    assert sigmas.shape == t_starts.shape, f"sigmas must have shape {tuple(t_starts.shape)}"
    counts = ops.counts_of(ray_indices)   # capacity-sized sample set (sampling(..., device_counts=True))
    offsets = ops.ray_offsets(ray_indices, n_rays) if counts is None else counts[0]
    n_dev = None if counts is None else counts[1]
    colors, opac, depth, weights, trans, alphas = ops.CompositeFunction.apply(
        t_starts, t_ends, sigmas, rgbs, offsets, n_rays, render_bkgd)
    extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs}
    io = sres.get("interal_output")
    if io is not None:
        if "latent_losses" in io:
            # reduce_along_rays(..., weights.detach(), "sum") (cednerf/render.py:105-113) as one segmented launch
            ridx = ray_indices.detach().to(torch.int64).contiguous()
            extras["latent_losses"] = ops.AccumulateFunction.apply(weights.detach(), io["latent_losses"], ridx, offsets,
                                                                   n_rays, n_dev)
        if "weight_losses" in io:
            wl = F.huber_loss(io["weight_losses"].float(), trans[:, None], reduction="none")
            extras["weight_losses"] = reduce_along_rays(ray_indices, wl * io["selector"][:, None], n_rays,
                                                        weights[:, None])
    return colors, opac, depth, extras
