"""Worker of tests/test_gpu_dp.py: launched once per GPU by torch.distributed.run.  Checks, on real devices,
dp.DistributedFusedAdam (fused reduce-scatter + Adam + all-gather over NVLink peer memory, csrc/dp.cu) against a
single-process run of the same step, and the occupancy replicas under dp.SharedRng.  Prints 'DP-OK' on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cednerf_b200 as cb  # noqa: E402
from cednerf_b200 import dp, workload  # noqa: E402


def loss_of(field, est, batch, cfg, rk):
    rays = cb.Rays(batch["origins"], batch["viewdirs"])
    rgb, acc, _, n_s, extra = cb.render_image(field, est, rays, render_bkgd=batch["color_bkgd"],
                                              timestamps=batch["timestamps"], jitter=batch["jitter"], **rk)
    assert n_s > 0
    return cb.losses.training_loss(rgb, acc, batch["pixels"], extra, acc_entropy_loss=True, weight_rgbper=True,
                                   use_feat_predict=bool(cfg.flags.get("use_feat_predict")))


def adam_reference(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-15):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    return p - (lr / bc1) * (m / (v.sqrt() / (bc2 ** 0.5) + eps)), m, v


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    cfg, n_rays = workload.TINY, 4096
    rk = workload.render_kwargs(cfg)
    est, field = workload.build_scene(cfg, dev, cb, seed=42)
    est.train(), field.train()
    gen = torch.Generator().manual_seed(5)
    full = {k: v.to(dev) for k, v in workload.draw_batch(cfg, n_rays * world, gen).items()}
    mine = {k: (v if v.dim() == 1 and v.numel() == 3 else v[rank * n_rays:(rank + 1) * n_rays].contiguous())
            for k, v in full.items()}
    params0 = [p.detach().clone() for p in field.parameters()]
    scale = 1024.0

    # ---- single-process reference on rank 0's device: the full batch, gradients only -----------------------------
    (loss_of(field, est, full, cfg, rk) * scale).backward()
    g_full = {n: p.grad.detach().clone() for n, p in field.named_parameters() if p.grad is not None}
    for p in field.parameters():
        p.grad = None

    # ---- data-parallel step -----------------------------------------------------------------------------------------
    provider = os.environ.get("CEDNERF_DP_PROVIDER", "auto")   # "auto": symmetric memory (NVLS) if the box has it; "ipc"
    opt = dp.DistributedFusedAdam(field.parameters(), lr=1e-2, eps=1e-15, provider=provider,
                                  nvls=True if provider == "auto" else "auto")   # exercise the in-switch path at N = 2 too
    opt.setup()
    if rank == 0:
        print("transport:", opt.transport(), flush=True)
    scaler = cb.optim.GradScaler(scale)
    table = field.hash_encoder.params
    assert table.data_ptr() == opt._p32.data_ptr(), "table parameter was not re-homed into peer memory"
    for it in range(2):
        opt.zero_grad()
        scaler.scale(loss_of(field, est, mine, cfg, rk)).backward()
        assert table.grad.data_ptr() == opt._g_table.data_ptr(), "autograd copied the peer gradient buffer"
        if it == 0:
            g_loc = {n: p.grad.detach().clone() for n, p in field.named_parameters() if p.grad is not None}
        scaler.step(opt)
        scaler.update()
        if it == 0:
            opt.sync_master()   # one-sided: the peers' owned ranges of the fp32 master (they cannot be a step ahead)
            p_after1 = {n: p.detach().clone() for n, p in field.named_parameters()}
    torch.cuda.synchronize()
    assert not opt.timed_out()
    assert float(scaler.get_scale()) == scale

    # (1) averaged local gradients == full-batch gradient (same loss, mean over rays)
    for n, g in g_loc.items():
        gs = g.clone()
        dist.all_reduce(gs)
        gs /= world
        err = float((gs - g_full[n]).norm() / g_full[n].norm().clamp_min(1e-30))
        # (fp16 gradient tensors inside the MLPs are rounded at half the magnitude in the full-batch run)
        assert err < 1e-3, (n, err)
        # (2) the fused step == Adam applied to that average (summed in rank order), first iteration
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g)
        acc = parts[0].clone()
        for q in parts[1:]:
            acc += q
        gavg = acc * ((1.0 / scale) / world)
        p0 = dict(zip([k for k, _ in field.named_parameters()], params0))[n]
        want, _, _ = adam_reference(p0, gavg, torch.zeros_like(p0), torch.zeros_like(p0), 1, 1e-2)
        got = p_after1[n]
        # Adam's first step is -lr * g / (|g| + eps): elements whose gradient is ~0 are ill-conditioned, compare the others
        solid = gavg.abs() > 1e-12
        d = (got - want)[solid].abs().max() if solid.any() else torch.zeros(())
        assert float(d) <= 2e-6, (n, float(d))
        assert torch.equal(got[~solid & (gavg == 0)], p0[~solid & (gavg == 0)])  # untouched entries did not move
    # (3) replicas are bit-identical across ranks: fp16 working copy, MLP parameters - and the fp32 master once every rank
    # has pulled the ranges it does not own (one-sided; the default all-gathers only the fp16 copy per step)
    if world > 1:
        own = table.detach().view(-1)[opt._lo:opt._hi].clone()
        opt.sync_master()
        assert torch.equal(own, table.detach().view(-1)[opt._lo:opt._hi])
        dist.barrier()
    for n, p in list(field.named_parameters()) + [("table_f16", field.hash_encoder.table_f16())]:
        if p.numel() == 0:
            continue
        parts = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(parts, p.detach().contiguous())
        for q in parts[1:]:
            assert torch.equal(parts[0], q), f"replicas of {n} differ"
    assert torch.equal(field.hash_encoder.table_f16().view(-1), table.detach().view(-1).half())
    assert field.hash_encoder.table_f16().data_ptr() == opt._p16.data_ptr()
    # (4) moments are sharded; gather_state reassembles them
    m_full, v_full = opt.gather_state()
    assert m_full.numel() == table.numel() and opt.state[table]["exp_avg"].numel() == opt._hi - opt._lo
    assert float(m_full.abs().sum()) > 0

    # (5) non-finite gradients anywhere skip the step everywhere and halve the scale
    before = table.detach().clone()
    opt.zero_grad()
    scaler.scale(loss_of(field, est, mine, cfg, rk)).backward()
    if rank == world - 1:
        field.mlp_head.params.grad[3] = float("inf")
    scaler.step(opt)
    scaler.update()
    torch.cuda.synchronize()
    assert torch.equal(before, table.detach()) and float(scaler.get_scale()) == scale / 2

    # (6) occupancy replicas: same draws on every rank -> identical occs / binaries (train_real.py:324-336)
    rng = dp.SharedRng(1234, dev)

    def occ_eval_fn(x):
        return field.query_density(x, rng.rand(x.shape[0], 1))["density"] * rk["render_step_size"]

    est2 = cb.OccGridEstimator(list(cfg.roi_aabb), resolution=cfg.occ_res, levels=cfg.occ_levels).to(dev).train()
    for step in (0, 16, 256, 272):
        est2.update_every_n_steps(step=step, occ_eval_fn=occ_eval_fn, occ_thre=1e-2, rng=rng)
    for t in (est2.occs, est2.binaries.to(torch.uint8)):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        for q in parts[1:]:
            assert torch.equal(parts[0], q), "occupancy replicas differ"
    assert bool(est2.binaries.any())
    dp.sync_occupancy(est2)

    opt.close()
    dist.barrier()
    if rank == 0:
        print("DP-OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
