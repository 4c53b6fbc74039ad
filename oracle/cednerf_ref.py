"""ORACLE (test infrastructure) — restatement of the reference's in-tree per-ray pipeline.

Follows, on CPU with the oracle's nerfacc/tcnn restatements underneath:
  SinusoidalEncoder / SinusoidalEncoderWithExp   cednerf/encoder.py:6-44, :46-90
  trunc_exp                                      cednerf/utils.py:27-43
  DNGPradianceField                              cednerf/model.py:97-488
  rendering                                      cednerf/render.py:58-176
  render_image / render_image_test               cednerf/utils.py:46-150, :153-318

PINNED: tests/golden/make_golden.py runs the reference's own files (imported from
/root/reference with nerfacc/tinycudann/taichi stubbed by oracle.*_ref) and freezes vectors that
tests/test_oracle_golden.py checks this restatement against.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import Optional

import torch
import torch.nn.functional as F

from . import nerfacc_ref as nf
from . import tcnn_ref as tc

Rays = namedtuple("Rays", ("origins", "viewdirs"))


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(max=15.0))


trunc_exp = _TruncExp.apply


def time_embed(t: torch.Tensor) -> torch.Tensor:
    """SinusoidalEncoder(1,0,4,True): [t, sin(2^k t) k=0..3, sin(2^k t + pi/2) k=0..3]."""
    tb = t.view(-1, 1) * torch.tensor([1.0, 2.0, 4.0, 8.0])
    return torch.cat([t.view(-1, 1), torch.sin(torch.cat([tb, tb + 0.5 * math.pi], -1))], -1)


def time_embed_attenuated(t: torch.Tensor, move_norm: torch.Tensor) -> torch.Tensor:
    """SinusoidalEncoderWithExp(1,0,4,True): [t, (sin 2^i t, cos 2^i t) * exp(-i 2^i |move|) i=0..3]."""
    sc = torch.tensor([1.0, 2.0, 4.0, 8.0])
    scm = torch.tensor([0.0, 2.0, 8.0, 24.0])
    tb = t.view(-1, 1) * sc
    att = torch.exp(-1 * (move_norm.view(-1, 1) * scm))
    s = torch.sin(tb) * att
    c = torch.sin(tb + 0.5 * math.pi) * att
    return torch.cat([t.view(-1, 1), torch.stack([s, c], -1).reshape(-1, 8)], -1)


class DNGPradianceField(torch.nn.Module):
    def __init__(self, aabb, num_dim=3, use_viewdirs=True, geo_feat_dim=15, base_resolution=16, n_levels=16,
                 n_features_per_level=2, dst_resolution=4096, log2_hashmap_size=19, use_feat_predict=False,
                 use_weight_predict=False, moving_step=1 / 4096, use_div_offsets=False, use_time_embedding=False,
                 use_time_attenuation=False, time_inject_before_sigma=True, seed=1337):
        super().__init__()
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=torch.float32))
        self.num_dim, self.use_viewdirs, self.geo_feat_dim = num_dim, use_viewdirs, geo_feat_dim
        self.use_feat_predict, self.use_weight_predict = use_feat_predict, use_weight_predict
        self.use_time_embedding, self.use_time_attenuation = use_time_embedding, use_time_attenuation
        self.time_inject_before_sigma, self.use_div_offsets = time_inject_before_sigma, use_div_offsets
        self.MOVING_STEP, self.loose_move = moving_step, False
        mlp = lambda h: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                         "n_neurons": 64, "n_hidden_layers": h}
        freq = {"otype": "Frequency", "n_frequencies": 4}
        self.xyz_wrap = tc.NetworkWithInputEncoding(4, 6 if use_div_offsets else 3, freq, mlp(3), seed + 1)
        self.direction_encoding = tc.Encoding(3, {"otype": "SphericalHarmonics", "degree": 2})
        b = math.exp(math.log(dst_resolution / base_resolution) / (n_levels - 1))
        self.hash_encoder = tc.Encoding(num_dim, {"otype": "HashGrid", "n_levels": n_levels,
                                                  "n_features_per_level": n_features_per_level,
                                                  "log2_hashmap_size": log2_hashmap_size,
                                                  "base_resolution": base_resolution, "per_level_scale": b}, seed + 2)
        base_in, self.geo_feat_dim_head = self.hash_encoder.n_output_dims, geo_feat_dim
        if use_time_embedding:
            # the reference's encoder modules carry these buffers into its checkpoints (cednerf/encoder.py:18-20, :55-60)
            self.time_encoder, self.time_encoder_feat = torch.nn.Module(), torch.nn.Module()
            self.time_encoder.register_buffer("scales", torch.tensor([2 ** i for i in range(4)]))
            self.time_encoder_feat.register_buffer("scales", torch.tensor([2 ** i for i in range(4)]))
            self.time_encoder_feat.register_buffer("scales_move", torch.tensor([i * 2 ** i for i in range(4)]))
            if time_inject_before_sigma:
                base_in += 9
            else:
                self.geo_feat_dim_head += 9
        self.mlp_base = tc.Network(base_in, 1 + geo_feat_dim, mlp(1), seed + 3)
        self.mlp_head = tc.Network((4 if use_viewdirs else 0) + self.geo_feat_dim_head, 3, mlp(2), seed + 4)
        if use_feat_predict:
            self.mlp_feat_prediction = tc.NetworkWithInputEncoding(4, self.hash_encoder.n_output_dims, freq, mlp(1), seed + 5)
        if use_weight_predict:
            self.mlp_weight_prediction = tc.NetworkWithInputEncoding(4, 1, freq, mlp(1), seed + 6)

    def query_move(self, x, t):
        off = self.xyz_wrap(torch.cat([x, t], -1))
        move = off[:, :3] * self.MOVING_STEP
        if self.use_div_offsets:
            move = move + torch.tanh(off[:, 3:]) * self.MOVING_STEP
        return x + move, move

    def query_density(self, x, t, return_feat=False, return_interal=False):
        if (not self.loose_move) and x.shape[0] > 0:
            x_move, move = self.query_move(x.view(-1, 3), t.view(-1, 1))
        else:
            x_move = x.view(-1, 3)
            move = torch.zeros_like(x_move[:, :1])
        lo, hi = self.aabb[:3], self.aabb[3:]
        x_move = (x_move - lo) / (hi - lo)
        selector = ((x_move > 0.0) & (x_move < 1.0)).all(-1)
        hash_feat = self.hash_encoder(x_move)
        feat = hash_feat
        time_encode = None
        if self.use_time_embedding:
            with torch.no_grad():
                if self.use_time_attenuation:
                    move = torch.linalg.norm(move.detach(), dim=-1)
                    time_encode = time_embed_attenuated(t.view(-1, 1), move.view(-1, 1))
                else:
                    time_encode = time_embed(t.view(-1, 1))
            if self.time_inject_before_sigma:
                feat = torch.cat([hash_feat, time_encode], -1)
        h = self.mlp_base(feat).float()
        raw, base_out = h[:, :1], h[:, 1:]
        res = {"density": trunc_exp(raw - 1) * selector[:, None]}
        if return_feat:
            res["base_mlp_out"] = (torch.cat([base_out, time_encode], -1)
                                   if self.use_time_embedding and not self.time_inject_before_sigma else base_out)
        if return_interal:
            io = {"move": move}
            if self.use_feat_predict or self.use_weight_predict:
                tf = torch.cat([x_move, t.view(-1, 1)], -1)
                io["selector"] = selector
                if self.use_feat_predict:
                    io["latent_losses"] = F.huber_loss(self.mlp_feat_prediction(tf), hash_feat,
                                                       reduction="none") * selector[:, None]
                if self.use_weight_predict:
                    io["weight_losses"] = self.mlp_weight_prediction(tf)
            res["interal_output"] = io
        return res

    def _query_rgb(self, dirs, embedding, apply_act=True):
        if self.use_viewdirs:
            dirs = dirs / torch.linalg.norm(dirs, dim=-1, keepdims=True)
            h = torch.cat([self.direction_encoding((dirs + 1.0) / 2.0), embedding.reshape(-1, self.geo_feat_dim_head)], -1)
        else:
            h = embedding.reshape(-1, self.geo_feat_dim_head)
        rgb = self.mlp_head(h).float()
        return torch.sigmoid(rgb) if apply_act else rgb

    def forward(self, positions, t, directions=None):
        res = self.query_density(positions, t, return_feat=True, return_interal=self.training)
        return self._query_rgb(directions, res["base_mlp_out"]), res


def reduce_along_rays(ray_indices, values, n_rays, weights=None, reduce="mean"):
    """cednerf/render.py:8-39 (scatter_reduce_ with include_self=True: 'mean' divides by count+1)."""
    src = values if weights is None else weights * values
    out = torch.zeros(n_rays, src.shape[-1], dtype=src.dtype)
    if ray_indices.numel() == 0:
        return out
    index = ray_indices[:, None].long().expand(-1, src.shape[-1])
    return out.scatter_reduce(0, index, src, reduce=reduce)


def rendering(t_starts, t_ends, ray_indices, n_rays, rgb_sigma_fn, render_bkgd=None):
    rgbs, sres = rgb_sigma_fn(t_starts, t_ends, ray_indices)
    sigmas = sres["density"].squeeze(-1)
    weights, trans, alphas = nf.render_weight_from_density(t_starts, t_ends, sigmas, ray_indices=ray_indices, n_rays=n_rays)
    extras = {"weights": weights, "alphas": alphas, "trans": trans, "sigmas": sigmas, "rgbs": rgbs}
    io = sres.get("interal_output")
    if io is not None:
        if "latent_losses" in io:
            extras["latent_losses"] = reduce_along_rays(ray_indices, io["latent_losses"], n_rays,
                                                        weights[:, None].detach(), "sum")
        if "weight_losses" in io:
            wl = F.huber_loss(io["weight_losses"].float(), trans[:, None], reduction="none")
            extras["weight_losses"] = reduce_along_rays(ray_indices, wl * io["selector"][:, None], n_rays, weights[:, None])
    colors = nf.accumulate_along_rays(weights, rgbs, ray_indices, n_rays)
    opac = nf.accumulate_along_rays(weights, None, ray_indices, n_rays)
    depth = nf.accumulate_along_rays(weights, (t_starts + t_ends)[:, None] / 2.0, ray_indices, n_rays)
    depth = depth / opac.clamp_min(torch.finfo(rgbs.dtype).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opac)
    return colors, opac, depth, extras


def _field_fns(field, rays, timestamps):
    def positions_of(t0, t1, ridx):
        o, d = rays.origins[ridx], rays.viewdirs[ridx]
        x = o + d * (t0 + t1)[:, None] / 2.0
        t = timestamps[ridx] if field.training else timestamps.expand_as(x[:, :1])
        return x, t, d

    def sigma_fn(t0, t1, ridx):
        x, t, _ = positions_of(t0, t1, ridx)
        return field.query_density(x, t)["density"].squeeze(-1)

    def rgb_sigma_fn(t0, t1, ridx):
        x, t, d = positions_of(t0, t1, ridx)
        return field(x, t, d)

    return sigma_fn, rgb_sigma_fn


def render_image(field, estimator, rays, near_plane=0.0, far_plane=1e10, render_step_size=1e-3, render_bkgd=None,
                 cone_angle=0.0, alpha_thre=0.0, test_chunk_size=8192, timestamps=None, jitter=None):
    shape = rays.origins.shape
    rays = Rays(rays.origins.reshape(-1, 3), rays.viewdirs.reshape(-1, 3))
    n = rays.origins.shape[0]
    chunk = n if field.training else test_chunk_size
    outs, infos = [], []
    for i in range(0, n, chunk):
        cr = Rays(rays.origins[i:i + chunk], rays.viewdirs[i:i + chunk])
        ts = timestamps[i:i + chunk] if (field.training and timestamps is not None) else timestamps
        sigma_fn, rgb_sigma_fn = _field_fns(field, cr, ts)
        ridx, t0, t1 = estimator.sampling(cr.origins, cr.viewdirs, sigma_fn=sigma_fn, near_plane=near_plane,
                                          far_plane=far_plane, render_step_size=render_step_size,
                                          stratified=field.training, cone_angle=cone_angle, alpha_thre=alpha_thre,
                                          jitter=None if jitter is None else jitter[i:i + chunk])
        rgb, opac, depth, extras = rendering(t0, t1, ridx, cr.origins.shape[0], rgb_sigma_fn, render_bkgd)
        extras.update(ray_indices=ridx, t_starts=t0, t_ends=t1)
        outs.append((rgb, opac, depth, len(t0)))
        infos.append(extras)
    rgb, opac, depth = (torch.cat([o[k] for o in outs], 0) for k in range(3))
    return (rgb.view(*shape[:-1], -1), opac.view(*shape[:-1], -1), depth.view(*shape[:-1], -1),
            sum(o[3] for o in outs), infos)


@torch.no_grad()
def render_image_test(max_samples, field, estimator, rays, near_plane=0.0, far_plane=1e10, render_step_size=1e-3,
                      render_bkgd=None, cone_angle=0.0, alpha_thre=0.0, early_stop_eps=1e-4, timestamps=None):
    shape = rays.origins.shape
    rays = Rays(rays.origins.reshape(-1, 3), rays.viewdirs.reshape(-1, 3))
    n = rays.origins.shape[0]
    _, rgb_sigma_fn = _field_fns(field, rays, timestamps)
    opacity, depth, rgb = torch.zeros(n, 1), torch.zeros(n, 1), torch.zeros(n, 3)
    alive = torch.ones(n, dtype=torch.bool)
    min_samples = 1 if cone_angle == 0 else 4
    near = torch.full((n,), float(near_plane))
    far = torch.full((n,), float(far_plane))
    t_mins, t_maxs, hits = nf.ray_aabb_intersect(rays.origins, rays.viewdirs, estimator.aabbs)
    t_sorted, t_indices = nf.sort_boundaries(t_mins, t_maxs)
    done = total = 0
    while done < max_samples:
        n_alive = int(alive.sum())
        if n_alive == 0:
            break
        k = max(min(n // n_alive, 64), min_samples)
        done += k
        iv, sm, term = nf.traverse_grids(rays.origins, rays.viewdirs, estimator.binaries, estimator.aabbs, near, far,
                                         render_step_size, cone_angle, k, True, alive, t_sorted, t_indices, hits)
        t0, t1 = iv.vals[iv.is_left], iv.vals[iv.is_right]
        ridx = sm.ray_indices[sm.is_valid]
        rgbs, sres = rgb_sigma_fn(t0, t1, ridx)
        w, _, _ = nf.render_weight_from_density(t0, t1, sres["density"].squeeze(-1), ray_indices=ridx, n_rays=n,
                                                prefix_trans=1 - opacity[ridx].squeeze(-1))
        nf.accumulate_along_rays_(w, rgbs, ridx, rgb)
        nf.accumulate_along_rays_(w, None, ridx, opacity)
        nf.accumulate_along_rays_(w, (t0 + t1)[:, None] / 2.0, ridx, depth)
        near = term
        alive = (opacity.view(-1) <= 1 - early_stop_eps) & (sm.packed_info[:, 1] == k)
        total += ridx.shape[0]
    rgb = rgb + render_bkgd * (1.0 - opacity)
    depth = depth / opacity.clamp_min(torch.finfo(torch.float32).eps)
    return rgb.view(*shape[:-1], -1), opacity.view(*shape[:-1], -1), depth.view(*shape[:-1], -1), total


def distortion(ray_ids, weights, t_starts, t_ends):
    """cednerf/losses.py:4-11: flatten_eff_distloss(w, mid-points, interval lengths, ray ids) of the third-party package
    torch_efficient_distloss (not vendored, version un-pinned in the reference; algorithm: Sun et al., "Improved Direct
    Voxel Grid Optimization", 2022, eq. 12-14 - the O(N) form of Mip-NeRF 360's distortion loss):
        L = [ sum_i d_i w_i^2 / 3 + 2 sum_i w_i (m_i W_i - M_i) ] / (ray_ids.max() + 1)
    with W_i / M_i the exclusive prefix sums of w and w m inside sample i's ray.  fp64; autograd gives d/dw."""
    w = weights.reshape(-1).double()
    t0, t1 = t_starts.reshape(-1).double(), t_ends.reshape(-1).double()
    if w.numel() == 0:
        return w.sum()
    m, d = (t0 + t1) / 2, t1 - t0
    n_rays = int(ray_ids.max()) + 1
    first = torch.ones_like(ray_ids, dtype=torch.bool)
    first[1:] = ray_ids[1:] != ray_ids[:-1]
    start = torch.cummax(torch.where(first, torch.arange(w.numel()), torch.zeros_like(ray_ids)), 0)[0]
    cw, cwm = torch.cumsum(w, 0), torch.cumsum(w * m, 0)
    base_w = torch.cat([w.new_zeros(1), cw])[start]      # cumulative sums just before the ray's first sample
    base_wm = torch.cat([w.new_zeros(1), cwm])[start]
    W, M = cw - w - base_w, cwm - w * m - base_wm
    return ((d * w * w / 3).sum() + (2 * w * (m * W - M)).sum()) / n_rays


def distortion_bruteforce(ray_ids, weights, t_starts, t_ends):
    """The definition the O(N) form is derived from (Barron et al., Mip-NeRF 360, eq. 15, per ray):
    sum_ij w_i w_j |m_i - m_j| + sum_i w_i^2 d_i / 3, summed over rays and divided by ray_ids.max() + 1."""
    w = weights.reshape(-1).double()
    m = ((t_starts.double() + t_ends.double()) / 2).reshape(-1)
    d = (t_ends.double() - t_starts.double()).reshape(-1)
    total = w.new_zeros(())
    for r in ray_ids.unique().tolist():
        k = ray_ids == r
        wr, mr, dr = w[k], m[k], d[k]
        total = total + (wr[:, None] * wr[None, :] * (mr[:, None] - mr[None, :]).abs()).sum() + (wr * wr * dr / 3).sum()
    return total / (int(ray_ids.max()) + 1)
