"""World-size-2 gloo checks of the one-process-per-GPU plumbing (cednerf_b200/dp.py) and of the synthetic workload's
host side.  No kernels run here: the data-parallel path only adds sharding and a gradient all-reduce around them."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cednerf_b200 import dp

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))  # same init on all ranks
    table = torch.nn.Parameter(torch.zeros(1000, 2))
    params = list(net.parameters()) + [table]
    reducer = dp.GradAllReducer(params, world)
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(64, 8, generator=g), torch.randn(64, 3, generator=g)
    idx_all = torch.randint(0, 1000, (64,), generator=g)
    lo, hi = dp.shard_range(64, rank, world)
    pred = net(x_all[lo:hi]) + table[idx_all[lo:hi]].sum(-1, keepdim=True)
    torch.nn.functional.mse_loss(pred, y_all[lo:hi]).backward()   # local mean over the local slice
    reducer.wait()                                                 # -> mean over ranks == global mean (equal slices)
    ms = dp.max_over_ranks(float(rank + 1), "cpu")
    tot = dp.sum_over_ranks(float(hi - lo), "cpu")
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in params], "ms": ms, "tot": tot}, out)
    dist.destroy_process_group()


def test_gradient_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "dp.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    table = torch.nn.Parameter(torch.zeros(1000, 2))
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(64, 8, generator=g), torch.randn(64, 3, generator=g)
    idx_all = torch.randint(0, 1000, (64,), generator=g)
    pred = net(x_all) + table[idx_all].sum(-1, keepdim=True)
    torch.nn.functional.mse_loss(pred, y_all).backward()
    for a, p in zip(got["grads"], list(net.parameters()) + [table]):
        torch.testing.assert_close(a, p.grad, rtol=1e-5, atol=1e-7)
    assert got["ms"] == 2.0 and got["tot"] == 64.0


class _TwoGradNode(torch.autograd.Function):
    """Stand-in for ops.FieldTrainFunction: ONE autograd node that produces the table gradient first and the other
    gradients afterwards, calling the table-gradient hook in between (what the fused backward does on the GPU)."""

    @staticmethod
    def forward(ctx, w, table, x, idx):
        ctx.save_for_backward(w, table, x, idx)
        return x @ w + table[idx].sum(-1, keepdim=True)

    @staticmethod
    def backward(ctx, g):
        from cednerf_b200 import ops

        w, table, x, idx = ctx.saved_tensors
        gt = torch.zeros_like(table).index_add_(0, idx, g.sum(-1, keepdim=True).expand(-1, 2).contiguous())
        if ops._table_grad_hook is not None:
            ops._table_grad_hook(gt)       # all-reduce starts here ...
        gw = x.t() @ g                      # ... and overlaps the rest of the node's work
        return gw, gt, None, None


def _worker_early(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cednerf_b200 import dp, ops

    g = torch.Generator().manual_seed(11)
    w = torch.nn.Parameter(torch.randn(8, 3, generator=g))
    table = torch.nn.Parameter(torch.randn(500, 2, generator=g))
    x_all, y_all = torch.randn(64, 8, generator=g), torch.randn(64, 3, generator=g)
    idx_all = torch.randint(0, 500, (64,), generator=g)
    reducer = dp.GradAllReducer([w, table], world)
    assert ops._table_grad_hook is not None
    lo, hi = dp.shard_range(64, rank, world)
    res = {}
    for it in range(2):  # second pass: .grad already exists, autograd accumulates into it and the early path stands down
        pred = _TwoGradNode.apply(w, table, x_all[lo:hi], idx_all[lo:hi])
        torch.nn.functional.mse_loss(pred, y_all[lo:hi]).backward()
        reducer.wait()
        res[f"w{it}"], res[f"table{it}"] = w.grad.clone(), table.grad.clone()
    reducer.remove()
    assert ops._table_grad_hook is None
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_table_gradient_allreduce_started_inside_the_fused_backward(tmp_path):
    out = str(tmp_path / "dp_early.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_early, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(11)
    w = torch.nn.Parameter(torch.randn(8, 3, generator=g))
    table = torch.nn.Parameter(torch.randn(500, 2, generator=g))
    x_all, y_all = torch.randn(64, 8, generator=g), torch.randn(64, 3, generator=g)
    idx_all = torch.randint(0, 500, (64,), generator=g)
    pred = _TwoGradNode.apply(w, table, x_all, idx_all)
    torch.nn.functional.mse_loss(pred, y_all).backward()
    for it in range(2):
        torch.testing.assert_close(got[f"w{it}"], w.grad * (it + 1), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(got[f"table{it}"], table.grad * (it + 1), rtol=1e-5, atol=1e-7)


def test_sharding_helpers():
    from cednerf_b200 import dp

    for n in (0, 1, 7, 300, 1014):
        for world in (1, 2, 3, 8):
            blocks = [dp.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1
            inter = sorted(i for r in range(world) for i in dp.shard_interleaved(n, r, world))
            assert inter == list(range(n))


def test_workload_host_side_is_seeded_and_shaped():
    from cednerf_b200 import workload as w

    cfg = w.DYNERF
    assert (cfg.width, cfg.height, cfg.n_frames, cfg.occ_levels, cfg.dst_resolution) == (1352, 1014, 300, 4, 8192)
    a = w.draw_batch(cfg, 512, torch.Generator().manual_seed(3))
    b = w.draw_batch(cfg, 512, torch.Generator().manual_seed(3))
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["origins"].shape == (512, 3) and a["timestamps"].shape == (512, 1)
    torch.testing.assert_close(a["viewdirs"].norm(dim=-1), torch.ones(512))
    assert float(a["timestamps"].min()) >= 0 and float(a["timestamps"].max()) <= 1
    occ = w.blob_occupancy(w.TINY)
    assert occ.shape == (2, 16, 16, 16) and 0.02 < float(occ[0].float().mean()) < 0.6
    # a coarser level is the OR-downsample of the finer one over the region they share
    fine = w.blob_occupancy(cfg)
    inner = fine[1][32:96, 32:96, 32:96]
    pooled = torch.nn.functional.max_pool3d(fine[0][None, None].float(), 2)[0, 0].bool()
    assert torch.equal(inner, pooled)
    o, d = w.frame_rays(w.TINY, 1, rows=(4, 9))
    assert o.shape == (5 * w.TINY.width, 3) and torch.equal(o[0], w.camera_centres(w.TINY)[1])
