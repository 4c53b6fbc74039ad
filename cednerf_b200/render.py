"""cednerf/render.py surface: rendering() (cednerf/render.py:58-176) and reduce_along_rays (:8-39)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def reduce_along_rays(ray_indices, values, n_rays=None, weights=None, reduce: str = "mean"):
    """cednerf/render.py:8-39: out = zeros(n_rays, C).scatter_reduce_(0, ray_indices, weights * values, reduce) with
    include_self=True ('mean' therefore divides a ray's sum by its sample count + 1).  'sum' and 'mean' - what the
    reference calls it with - are one segmented launch forward and one backward (cednerf_accumulate_fwd / _bwd, gradients
    to both `values` and `weights`); ray_indices sorted, as the sampler emits them."""
    assert ray_indices.dim() == 1 and values.dim() == 2
    if not values.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if reduce not in ("sum", "mean"):
        raise NotImplementedError("reduce_along_rays: 'sum' and 'mean' (what the reference uses)")
    if weights is not None:
        assert values.shape[0] == weights.shape[0], f"Invalid shapes: {tuple(values.shape)} vs {tuple(weights.shape)}"
    if ray_indices.numel() == 0:
        assert n_rays is not None
        return torch.zeros(n_rays, values.shape[-1], device=values.device)
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1
    counts = ops.counts_of(ray_indices)
    offsets = ops.ray_offsets(ray_indices, n_rays) if counts is None else counts[0]
    n_dev = None if counts is None else counts[1]
    ridx = ray_indices.detach().to(torch.int64).contiguous()
    w = torch.ones(values.shape[0], device=values.device) if weights is None else weights.reshape(-1)
    out = ops.AccumulateFunction.apply(w, values, ridx, offsets, n_rays, n_dev)
    if reduce == "mean":
        out = out / (offsets[1:] - offsets[:-1] + 1).to(out.dtype)[:, None]
    return out


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
                                                        weights[:, None])   # 'mean': segmented kernel + count division
    return colors, opac, depth, extras
