"""Full-size checks at BASELINE.json's sizes (DyNeRF-shaped: 2^18 rays, 4 x 128^3 occupancy levels, 16-level 2^21 hash
table): size-independent properties over the whole batch plus bit-exact / tolerance comparison against the oracle on a
random sub-sample the oracle finishes in seconds."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nerfacc_ref as nf  # noqa: E402
from oracle import tcnn_ref as tc  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def scene():
    import cednerf_b200 as cb
    from cednerf_b200 import workload as w

    cfg = w.DYNERF
    est, field = w.build_scene(cfg, DEV, cb, seed=42)
    batch = w.draw_batch(cfg, 2 ** 18, torch.Generator().manual_seed(1000))
    return cb, w, cfg, est.train(), field.train(), batch


def test_march_full_batch_properties_and_oracle_subsample(scene):
    cb, w, cfg, est, field, batch = scene
    o, d, jit = batch["origins"].to(DEV), batch["viewdirs"].to(DEV), batch["jitter"]
    ridx, t0, t1, packed = est.march(o, d, cfg.near_plane, 1e10, cfg.render_step_size, cfg.cone_angle, True, jit)
    n = o.shape[0]
    assert ridx.numel() > 4 * n                                           # a realistic sample load
    assert int(packed[:, 1].sum()) == ridx.numel() and int(packed[-1].sum()) == ridx.numel()
    assert bool((ridx[1:] >= ridx[:-1]).all())                            # packed by ray
    same = ridx[1:] == ridx[:-1]
    assert bool((t0[1:][same] >= t1[:-1][same]).all())                    # ... then by t, intervals do not overlap
    assert bool((t1 > t0).all()) and bool((t0 >= cfg.near_plane).all())
    assert torch.equal(torch.bincount(ridx, minlength=n), packed[:, 1])
    # every sample midpoint lies in an occupied cell of the finest level that contains it
    x = o[ridx] + d[ridx] * ((t0 + t1) * 0.5)[:, None]
    lvl = torch.clamp(torch.ceil(torch.log2(x.abs().max(-1).values.clamp_min(1e-9))), min=0).long()
    inside = lvl < cfg.occ_levels
    assert float(inside.float().mean()) > 0.999
    half = (2.0 ** lvl.float())[:, None]
    cell = torch.clamp(((x + half) / (2 * half) * cfg.occ_res).long(), 0, cfg.occ_res - 1)
    occ = est.binaries[lvl.clamp(max=cfg.occ_levels - 1), cell[:, 0], cell[:, 1], cell[:, 2]]
    assert float(occ[inside].float().mean()) > 0.995                      # midpoints on a cell face may round across
    # oracle on a random 4096-ray sub-sample: identical counts, indices and t values
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(5))[:4096].sort().values
    near = torch.full((4096,), cfg.near_plane) + jit[sel] * cfg.render_step_size
    r_ref, a_ref, b_ref, p_ref, _ = nf.traverse_grids(batch["origins"][sel], batch["viewdirs"][sel], est.binaries.cpu(),
                                                      est.aabbs.cpu(), near, torch.full((4096,), 1e10),
                                                      cfg.render_step_size, cfg.cone_angle, packed_only=True)
    assert torch.equal(packed[sel.to(DEV), 1].cpu(), p_ref[:, 1])
    keep = torch.isin(ridx, sel.to(DEV))
    assert torch.equal(t0[keep].cpu(), a_ref) and torch.equal(t1[keep].cpu(), b_ref)


def test_hash_table_full_size_linearity_and_oracle_subsample(scene):
    cb, w, cfg, est, field, batch = scene
    enc = field.hash_encoder
    assert enc.params.numel() == 2 * 23928800
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2 ** 16, 3, generator=g).to(DEV)
    with torch.no_grad():
        y1 = enc(x).float()
        saved = enc.params.detach().clone()
        enc.params.mul_(0.5)                      # exact in fp16/fp32: the encoding is linear in the table
        y2 = enc(x).float()
        enc.params.copy_(saved)
    torch.testing.assert_close(y2 * 2, y1, rtol=2e-3, atol=1e-3)
    levels = tc.grid_levels(16, 16, math.log(enc.cfg["per_level_scale"]), 2 ** 21)
    y_ref = tc.hashgrid_forward(x[:4096].cpu(), saved.cpu().view(-1, 2), levels)
    assert torch.equal(y1[:4096].cpu(), y_ref)


def test_composite_full_batch_properties(scene):
    cb, w, cfg, est, field, batch = scene
    n_rays = 2 ** 16
    g = torch.Generator().manual_seed(9)
    cnt = torch.poisson(torch.full((n_rays,), 16.0), generator=g).long()
    ridx = torch.repeat_interleave(torch.arange(n_rays), cnt).to(DEV)
    s = ridx.numel()
    t0 = torch.rand(s, generator=g).to(DEV)
    t1 = t0 + 0.01
    sig = (torch.rand(s, generator=g) * 80).to(DEV)
    rgb = torch.rand(s, 3, generator=g).to(DEV)
    off = cb.ops.ray_offsets(ridx, n_rays)
    colors, opac, depth, wts, tr, al = cb.ops.CompositeFunction.apply(t0, t1, sig, rgb, off, n_rays, None)
    assert bool((opac <= 1 + 1e-5).all()) and bool((opac >= 0).all())     # sum of weights <= 1
    same = ridx[1:] == ridx[:-1]
    assert bool((tr[1:][same] <= tr[:-1][same] + 1e-7).all())             # transmittance is monotone along a ray
    torch.testing.assert_close(opac.view(-1), torch.zeros(n_rays, device=DEV).index_add_(0, ridx, wts), rtol=1e-4, atol=1e-5)
    assert bool((colors <= 1 + 1e-4).all())
    # first sample of every ray has T = 1
    first = off[:-1][cnt.to(DEV) > 0]
    assert bool((tr[first] == 1).all())


def test_full_train_step_and_frame_render(scene):
    cb, w, cfg, est, field, batch = scene
    b = {k: v.to(DEV) for k, v in batch.items()}
    rk = w.render_kwargs(cfg)
    rays = cb.Rays(b["origins"], b["viewdirs"])
    rgb, acc, depth, n_s, extra = cb.render_image(field, est, rays, render_bkgd=b["color_bkgd"],
                                                  timestamps=b["timestamps"], jitter=b["jitter"], **rk)
    assert 2 ** 19 < n_s < 2 ** 22                                         # ~2^20 surviving samples for 2^18 rays
    loss = torch.nn.functional.mse_loss(rgb, b["pixels"]) + extra[0]["latent_losses"].mean()
    field.zero_grad()
    (loss * 1024).backward()
    assert bool(torch.isfinite(loss))
    for k, p in field.named_parameters():
        if p.numel():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    assert float(field.hash_encoder.params.grad.abs().sum()) > 0 and float(field.xyz_wrap.params.grad.abs().sum()) > 0
    # one full 1352 x 1014 frame through the eval marcher
    field.eval(), est.eval()
    o, d = w.frame_rays(cfg, 0)
    frame = cb.Rays(o.to(DEV).view(cfg.height, cfg.width, 3), d.to(DEV).view(cfg.height, cfg.width, 3))
    img, a, dep, n_tot = cb.render_image_test(1024, field, est, frame, render_bkgd=torch.zeros(3, device=DEV),
                                              timestamps=torch.tensor([[0.5]], device=DEV), **rk)
    field.train(), est.train()
    assert img.shape == (cfg.height, cfg.width, 3) and n_tot > 10 ** 6
    assert bool(torch.isfinite(img).all()) and float(a.max()) <= 1 + 1e-4 and float(a.max()) > 0.5
