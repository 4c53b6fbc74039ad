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
