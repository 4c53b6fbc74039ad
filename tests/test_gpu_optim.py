"""GPU parity of the fused optimiser step against the reference's own fallback optimiser,
torch.optim.Adam(lr, eps=1e-15) under torch's GradScaler(2**10) (train_real.py:252, 267-274, 412-420), run on the CPU in
fp32 on identical gradients.  Tolerance: parameters and Adam moments rtol 2e-5 / atol 1e-7 after 6 steps (different
but equivalent fp32 operation order: lerp vs b1*m+(1-b1)*g), fp16 copy bit-equal to p.half(), loss scale equal."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(p0s, grads, scale0, lr, eps, inf_at):
    """torch.optim.Adam + the GradScaler protocol restated on the CPU: unscale, skip on non-finite, backoff / growth."""
    ps = [torch.nn.Parameter(p.clone()) for p in p0s]
    opt = torch.optim.Adam(ps, lr=lr, eps=eps)
    scale, growth_tracker = float(scale0), 0
    for it, gs in enumerate(grads):
        found = any(not torch.isfinite(g).all() for g in gs)
        if not found:
            for p, g in zip(ps, gs):
                p.grad = g / scale
            opt.step()
            growth_tracker += 1
            if growth_tracker == 2000:
                scale, growth_tracker = scale * 2.0, 0
        else:
            scale, growth_tracker = scale * 0.5, 0
    return ps, opt, scale


@pytest.mark.parametrize("stock_scaler", [False, True])
def test_fused_adam_and_scaler_match_torch_adam(stock_scaler):
    import cednerf_b200 as cb
    from cednerf_b200 import ops

    g = torch.Generator().manual_seed(7)
    sizes = [3 * 4096 + 37, 100, 4096 * 2, 0]
    p0s = [(torch.rand(n, generator=g) - 0.5) * 2e-2 for n in sizes]
    scale0, lr, eps, steps, inf_at = 2.0 ** 10, 1e-2, 1e-15, 6, 3
    grads = []
    for it in range(steps):
        gs = [torch.randn(n, generator=g) * 1e-3 * scale0 for n in sizes]
        if it == inf_at:
            gs[0][5] = float("inf")
            gs[2][77] = float("nan")
        grads.append(gs)
    ref_ps, ref_opt, ref_scale = _reference(p0s, grads, scale0, lr, eps, inf_at)

    ps = [torch.nn.Parameter(p.clone().to(DEV)) for p in p0s]
    cache = ops._F16Cache()
    ps[0]._cednerf_f16 = cache  # what tcnn.Encoding(HashGrid) does for its table
    opt = cb.optim.FusedAdam(ps, lr=lr, eps=eps)
    scaler = torch.amp.GradScaler("cuda", init_scale=scale0) if stock_scaler else cb.optim.GradScaler(scale0)
    launches0 = cb._lib.launch_count()
    for it, gs in enumerate(grads):
        opt.zero_grad()
        # the protocol of train_real.py:412-420; the scaled loss's gradient is substituted by the shared draw
        scaler.scale(torch.zeros((), device=DEV))
        for p, gr in zip(ps, gs):
            p.grad = gr.to(DEV)
        if it == inf_at:  # a parity scale mismatch would otherwise hide here: 2^10 before this step
            assert scaler.get_scale() == scale0
        scaler.step(opt)
        scaler.update()
    assert cb._lib.launch_count() > launches0
    assert scaler.get_scale() == ref_scale == scale0 / 2
    for p, q in zip(ps, ref_ps):
        torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=2e-5, atol=1e-7)
        if p.numel():
            st, rt = opt.state[p], ref_opt.state[q]
            torch.testing.assert_close(st["exp_avg"].cpu(), rt["exp_avg"], rtol=2e-5, atol=1e-9)
            torch.testing.assert_close(st["exp_avg_sq"].cpu(), rt["exp_avg_sq"], rtol=2e-5, atol=1e-12)
            assert float(st["step"].item()) == float(rt["step"]) == steps - 1
    # fp16 working copy: written by the update pass, bit-equal to a cast, and already keyed to the new version

    def no_recast(_):
        raise AssertionError("the fp16 copy should not be recast after a fused step")

    assert torch.equal(cache.get(ps[0], no_recast), ps[0].detach().half())


def test_fused_adam_without_scaler_and_weight_decay_modes():
    import cednerf_b200 as cb

    g = torch.Generator().manual_seed(3)
    p0 = (torch.rand(5000, generator=g) - 0.5)
    gr = [torch.randn(5000, generator=g) * 0.1 for _ in range(4)]
    for adam_w, ref_cls in ((False, torch.optim.Adam), (True, torch.optim.AdamW)):
        q = torch.nn.Parameter(p0.clone())
        ropt = ref_cls([q], lr=3e-3, eps=1e-8, weight_decay=1e-2)
        p = torch.nn.Parameter(p0.clone().to(DEV))
        opt = cb.optim.FusedAdam([p], lr=3e-3, eps=1e-8, weight_decay=1e-2, adam_w_mode=adam_w)
        for x in gr:
            q.grad, p.grad = x.clone(), x.to(DEV)
            ropt.step()
            opt.step()
        torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=2e-5, atol=1e-7)


def test_fused_adam_resumes_from_a_state_dict():
    """Checkpoint / resume (train_real.py:433-441 saves the optimiser state): 3 steps, state_dict -> new optimiser, 2 more
    steps == 5 uninterrupted steps (bias corrections continue from the loaded step count)."""
    import cednerf_b200 as cb

    g = torch.Generator().manual_seed(9)
    p0 = (torch.rand(9000, generator=g) - 0.5)
    grads = [torch.randn(9000, generator=g) * 0.1 for _ in range(5)]

    def run(opt, p, gs):
        for x in gs:
            p.grad = x.to(DEV)
            opt.step()

    a = torch.nn.Parameter(p0.clone().to(DEV))
    opt_a = cb.optim.FusedAdam([a], lr=3e-3, eps=1e-15)
    run(opt_a, a, grads)
    b = torch.nn.Parameter(p0.clone().to(DEV))
    opt_b = cb.optim.FusedAdam([b], lr=3e-3, eps=1e-15)
    run(opt_b, b, grads[:3])
    sd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt_b.state_dict().items()}
    c = torch.nn.Parameter(b.detach().clone())
    opt_c = cb.optim.FusedAdam([c], lr=3e-3, eps=1e-15)
    opt_c.load_state_dict(sd)
    run(opt_c, c, grads[3:])
    assert torch.equal(c.detach(), a.detach())
    assert float(opt_c.state[c]["step"].item()) == 5.0


def test_fused_adam_more_than_eight_tensors_and_per_group_lr():
    """More parameter tensors than one launch descriptor holds (8): the step counter must advance once, and every group
    keeps its own learning rate."""
    import cednerf_b200 as cb

    g = torch.Generator().manual_seed(21)
    sizes = [33, 4096 * 3, 7, 5000, 64, 4097, 1, 900, 12288, 77, 4096]
    p0s = [torch.rand(n, generator=g) - 0.5 for n in sizes]
    grads = [[torch.randn(n, generator=g) * 0.05 for n in sizes] for _ in range(3)]
    ref = [torch.nn.Parameter(p.clone()) for p in p0s]
    ours = [torch.nn.Parameter(p.clone().to(DEV)) for p in p0s]
    groups = lambda ps: [{"params": ps[:5], "lr": 1e-2}, {"params": ps[5:], "lr": 3e-4}]  # noqa: E731
    ropt = torch.optim.Adam(groups(ref), eps=1e-15)
    opt = cb.optim.FusedAdam(groups(ours), eps=1e-15)
    for gs in grads:
        for q, p, x in zip(ref, ours, gs):
            q.grad, p.grad = x.clone(), x.to(DEV)
        ropt.step()
        opt.step()
    assert float(opt.state[ours[0]]["step"].item()) == 3.0
    for q, p in zip(ref, ours):
        torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=2e-5, atol=1e-7)
