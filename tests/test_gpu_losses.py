"""GPU parity of the fused training loss (csrc/losses.cu) against the reference's own formulation in plain PyTorch fp32
(train_real.py:369-409, canonical flags -ae -wr -f): value rtol 1e-5, gradients rtol 1e-4 / atol 1e-9."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_loss(rgb, acc, pixels, ex, ae=True, wr=True, f=True):
    loss = torch.nn.functional.mse_loss(rgb, pixels)
    extra = 0.0
    if ae:
        t_last = (1 - acc).clamp(1e-6, 1 - 1e-6)
        extra = extra + (-(t_last * torch.log(t_last) + (1 - t_last) * torch.log(1 - t_last)).mean()) * 1e-3
    if wr:
        rgbper = (ex["rgbs"] - pixels[ex["ray_indices"]]).pow(2).sum(dim=-1)
        extra = extra + (rgbper * ex["weights"].detach()).sum() / pixels.shape[0] * 1e-3
    if f:
        extra = extra + ex["latent_losses"].mean()
    return loss + extra


@pytest.mark.parametrize("flags", [(True, True, True), (False, True, False), (True, False, True), (False, False, False)])
@pytest.mark.parametrize("n_rays,n_samples", [(1000, 3777), (70001, 300123), (5, 0)])
def test_training_loss_matches_pytorch(flags, n_rays, n_samples):
    import cednerf_b200 as cb

    g = torch.Generator().manual_seed(n_rays + n_samples)
    mk = lambda *s: torch.rand(*s, generator=g).to(DEV)  # noqa: E731
    pixels = mk(n_rays, 3)
    acc0 = mk(n_rays, 1)
    acc0[:3] = torch.tensor([[0.0], [1.0], [1.0 - 1e-7]], device=DEV)[: min(3, n_rays)]   # the clamp's flat ends
    ridx = torch.sort(torch.randint(0, n_rays, (n_samples,), generator=g))[0].to(DEV)
    leaves = [mk(n_rays, 3), acc0, mk(n_samples, 3), mk(n_rays, 32) * 0.1]
    weights = mk(n_samples)
    outs = []
    for impl in ("ref", "ours"):
        rgb, acc, rgbs, lat = (t.clone().requires_grad_(True) for t in leaves)
        ex = {"rgbs": rgbs, "weights": weights, "ray_indices": ridx, "latent_losses": lat}
        if impl == "ref":
            loss = reference_loss(rgb, acc, pixels, ex, *flags)
        else:
            launches = cb._lib.launch_count()
            loss = cb.losses.training_loss(rgb, acc, pixels, [ex], *flags)
            assert cb._lib.launch_count() == launches + 2
        (loss * 1024.0).backward()
        outs.append((loss.detach(), [None if t.grad is None else t.grad.clone() for t in (rgb, acc, rgbs, lat)]))
    (l_ref, g_ref), (l_our, g_our) = outs
    torch.testing.assert_close(l_our, l_ref, rtol=1e-5, atol=1e-8)
    for a, b in zip(g_our, g_ref):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0
        else:
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-9)


def test_distortion_loss_and_ray_generation_against_the_oracle():
    import cednerf_b200 as cb
    from oracle import cednerf_ref as cr

    g = torch.Generator().manual_seed(8)
    n_rays = 5000
    counts = torch.randint(0, 70, (n_rays,), generator=g)
    counts[::7] = 0
    counts[-5:] = 0
    ray_ids = torch.repeat_interleave(torch.arange(n_rays), counts)
    n = ray_ids.numel()
    dt = torch.rand(n, generator=g) * 0.02 + 1e-3
    csum = torch.cumsum(dt, 0)
    first = torch.ones(n, dtype=torch.bool)
    first[1:] = ray_ids[1:] != ray_ids[:-1]
    start = torch.cummax(torch.where(first, torch.arange(n), torch.zeros(n, dtype=torch.long)), 0)[0]
    t0 = 0.2 + csum - dt - (csum - dt)[start]
    t1 = t0 + dt
    w = torch.rand(n, generator=g) * 0.2
    wr = w.clone().double().requires_grad_(True)
    want = cr.distortion(ray_ids, wr, t0, t1)
    (gwant,) = torch.autograd.grad(want, wr)
    wg = w.to(DEV).requires_grad_(True)
    got = cb.losses.distortion(ray_ids.to(DEV), wg, t0.to(DEV), t1.to(DEV), n_rays=n_rays)
    (ggot,) = torch.autograd.grad(got * 3.0, wg)
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    torch.testing.assert_close(ggot.cpu().double() / 3.0, gwant, rtol=1e-4, atol=1e-9)

    # ray generation: the reference's element-wise formulation (dnerf_3d_video_IS.py:339-358), both conventions
    from cednerf_b200 import workload as w_
    for opengl, cfg in ((False, w_.TINY), (True, w_.DNERF)):
        K = torch.tensor([[cfg.focal, 0, cfg.width / 2], [0, cfg.focal, cfg.height / 2], [0, 0, 1.0]])
        poses = w_.spiral_poses(w_.TINY, 4) if not opengl else torch.stack([w_.orbit_pose(4.0, 0.2 * k) for k in range(4)])
        x = torch.randint(0, cfg.width, (3000,), generator=g)
        y = torch.randint(0, cfg.height, (3000,), generator=g)
        iid = torch.randint(0, 4, (3000,), generator=g)
        s = -1.0 if opengl else 1.0
        cam = torch.nn.functional.pad(torch.stack([(x - K[0, 2] + 0.5) / K[0, 0], (y - K[1, 2] + 0.5) / K[1, 1] * s], -1), (0, 1), value=s)
        c2w = poses[iid]
        dirs = (cam[:, None, :] * c2w[:, :3, :3]).sum(-1)
        want_d = dirs / torch.linalg.norm(dirs, dim=-1, keepdims=True)
        rays, raw = cb.utils.generate_rays(K, poses.to(DEV), cfg.width, cfg.height, opengl, x=x, y=y, image_id=iid,
                                           return_directions=True)
        torch.testing.assert_close(rays.viewdirs.cpu(), want_d, rtol=0, atol=2e-7)
        torch.testing.assert_close(raw.cpu(), dirs, rtol=0, atol=2e-7)
        assert torch.equal(rays.origins.cpu(), c2w[:, :3, 3])
        frame = cb.utils.generate_rays(K, poses[1].to(DEV), cfg.width, cfg.height, opengl)
        assert frame.origins.shape == (cfg.height, cfg.width, 3)
        o_ref, d_ref = w_.pose_rays(cfg, poses[1], opengl)
        torch.testing.assert_close(frame.viewdirs.view(-1, 3).cpu(), d_ref, rtol=0, atol=3e-7)
