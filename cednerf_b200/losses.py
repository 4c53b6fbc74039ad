"""Loss of the reference's training loop for its canonical flags (train_real.py:369-409), SURVEY.md §8f N4:

    loss = F.mse_loss(rgb, pixels)
         + 1e-3 * entropy(clamp(1 - acc, 1e-6, 1 - 1e-6)).mean()                                   (-ae)
         + 1e-3 * ((rgbs - pixels[ray_indices])**2).sum(-1) @ weights.detach() / n_rays             (-wr)
         + latent_losses.mean()                                                                     (-f)

as one forward and one backward launch (`csrc/losses.cu`) instead of ~50 element-wise / reduction launches and autograd
nodes.  `training_loss(rgb, acc, pixels, extras, ...)` takes what `render_image` returns."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream
from .ops import _f32c, _sempty, counts_of

F32, F64, I64 = torch.float32, torch.float64, torch.int64


class TrainingLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, acc, pixels, rgbs, weights, ray_indices, latent, w_entropy, w_rgbper, n_dev=None):
        rgb, pixels = _f32c(rgb), _f32c(pixels)
        n_rays, dev = rgb.shape[0], rgb.device
        acc_c = None if acc is None else _f32c(acc).view(-1)
        rgbs_c = None if rgbs is None else _f32c(rgbs)
        w_c = None if rgbs is None else _f32c(weights)
        ridx = None if rgbs is None else ray_indices.detach().to(I64).contiguous()
        lat = None if latent is None else _f32c(latent)
        n_s = 0 if rgbs_c is None else rgbs_c.shape[0]
        n_lat = 0 if lat is None else lat.shape[-1]
        sums = torch.empty(4, dtype=F64, device=dev)
        loss = torch.empty(1, dtype=F32, device=dev)
        call("cednerf_training_loss_fwd", ptr(rgb), ptr(acc_c), ptr(pixels), n_rays, ptr(rgbs_c), ptr(w_c), ptr(ridx), n_s,
             ptr(lat), n_lat, float(w_entropy), float(w_rgbper), ptr(sums), ptr(loss), ptr(n_dev), stream())
        keep = [t if t is not None else rgb for t in (acc_c, rgbs_c, w_c, ridx)]
        ctx.save_for_backward(rgb, pixels, *keep)
        ctx.flags = (acc_c is not None, rgbs_c is not None, n_lat, n_s, float(w_entropy), float(w_rgbper))
        ctx.shapes = (None if acc is None else acc.shape, None if latent is None else latent.shape)
        ctx.n_dev = n_dev
        ctx.set_materialize_grads(False)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return (None,) * 10
        rgb, pixels, acc_c, rgbs_c, w_c, ridx = ctx.saved_tensors
        has_acc, has_rgbs, n_lat, n_s, w_entropy, w_rgbper = ctx.flags
        n_rays, dev = rgb.shape[0], rgb.device
        need = ctx.needs_input_grad
        g = g.detach().to(F32).reshape(1).contiguous()
        d_rgb = torch.empty_like(rgb) if need[0] else None
        d_acc = torch.empty(n_rays, device=dev) if (has_acc and need[1]) else None
        d_rgbs = _sempty(rgbs_c.shape[0], 3, device=dev) if (has_rgbs and need[3]) else None
        d_lat = torch.empty(n_rays, n_lat, device=dev) if (n_lat and need[6]) else None
        call("cednerf_training_loss_bwd", ptr(g), ptr(rgb), ptr(acc_c) if has_acc else None, ptr(pixels), n_rays,
             ptr(rgbs_c) if has_rgbs else None, ptr(w_c) if has_rgbs else None, ptr(ridx) if has_rgbs else None, n_s,
             n_lat, w_entropy, w_rgbper, ptr(d_rgb), ptr(d_acc), ptr(d_rgbs), ptr(d_lat), ptr(ctx.n_dev), stream())
        acc_shape, lat_shape = ctx.shapes
        return (d_rgb, None if d_acc is None else d_acc.view(acc_shape), None, d_rgbs, None, None,
                None if d_lat is None else d_lat.view(lat_shape), None, None, None)


def training_loss(rgb, acc, pixels, extras, acc_entropy_loss: bool = True, weight_rgbper: bool = True,
                  use_feat_predict: bool = True, distortion_loss: bool = False) -> torch.Tensor:
    """`extras` is the list `render_image` returns (one dict per ray chunk; training renders the batch as one chunk).
    distortion_loss: + 1e-3 * distortion(ray_indices, weights, t_starts, t_ends) (the `-d` flag, train_real.py:380-386)."""
    if len(extras) != 1:
        raise ValueError("training_loss fuses the one-chunk case of a training batch (render_image in training mode)")
    ex = extras[0]
    latent = ex.get("latent_losses") if use_feat_predict else None
    counts = counts_of(ex["ray_indices"])   # capacity-sized sample set: the live count stays on the device
    if distortion_loss:
        return _fused_terms(rgb, acc, pixels, ex, latent, counts, acc_entropy_loss, weight_rgbper) + 1e-3 * distortion(
            ex["ray_indices"], ex["weights"], ex["t_starts"], ex["t_ends"], n_rays=rgb.shape[0])
    return TrainingLossFunction.apply(rgb, acc if acc_entropy_loss else None, pixels,
                                      ex["rgbs"] if weight_rgbper else None, ex["weights"].detach(), ex["ray_indices"],
                                      latent, 1e-3, 1e-3, None if counts is None else counts[1])


def _fused_terms(rgb, acc, pixels, ex, latent, counts, acc_entropy_loss, weight_rgbper):
    return TrainingLossFunction.apply(rgb, acc if acc_entropy_loss else None, pixels,
                                      ex["rgbs"] if weight_rgbper else None, ex["weights"].detach(), ex["ray_indices"],
                                      latent, 1e-3, 1e-3, None if counts is None else counts[1])


class DistortionFunction(torch.autograd.Function):
    """flatten_eff_distloss(weights, mid-points, interval lengths, ray ids) of packed samples; gradient w.r.t. weights."""

    @staticmethod
    def forward(ctx, weights, t_starts, t_ends, offsets, n_rays):
        w, t0, t1 = _f32c(weights).view(-1), _f32c(t_starts).view(-1), _f32c(t_ends).view(-1)
        dev = w.device
        work = torch.empty(int(_lib.load().cednerf_distortion_workspace_bytes()) // 8, dtype=F64, device=dev)
        loss = torch.empty(1, dtype=F32, device=dev)
        inv = torch.empty(1, dtype=F32, device=dev)
        call("cednerf_distortion_fwd", ptr(w), ptr(t0), ptr(t1), ptr(offsets), int(n_rays), ptr(work), ptr(loss), ptr(inv),
             stream())
        ctx.save_for_backward(w, t0, t1, offsets, inv)
        ctx.n_rays, ctx.shape = int(n_rays), weights.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        w, t0, t1, offsets, inv = ctx.saved_tensors
        g = g.detach().to(F32).reshape(1).contiguous()
        gw = torch.zeros_like(w)
        call("cednerf_distortion_bwd", ptr(w), ptr(t0), ptr(t1), ptr(offsets), ctx.n_rays, ptr(g), ptr(inv), ptr(gw), stream())
        return gw.view(ctx.shape), None, None, None, None


def distortion(ray_ids, weights, t_starts, t_ends, n_rays=None):
    """cednerf/losses.py:4-11 (the `-d` flag of the HyperNeRF recipe, train_real.py:380-386): the distortion regulariser
    of Mip-NeRF 360 in its O(N) form (torch_efficient_distloss.flatten_eff_distloss).  `ray_ids` sorted, as the sampler
    emits them.  n_rays: number of rays of the batch (only sizes the per-ray launch; the loss is normalised by
    ray_ids.max() + 1 as in the package).  With capacity-sized inputs (ops.counts_of) the live prefix is used."""
    from . import ops

    counts = counts_of(ray_ids)
    if counts is not None:
        offsets = counts[0]
        n_rays = offsets.numel() - 1
    else:
        if n_rays is None:
            n_rays = int(ray_ids.max()) + 1 if ray_ids.numel() else 0
        offsets = ops.ray_offsets(ray_ids, n_rays)
    return DistortionFunction.apply(weights, t_starts, t_ends, offsets, n_rays)
