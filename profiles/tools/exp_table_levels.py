#!/usr/bin/env python
"""Diagnostic (GPU): time cednerf_hashgrid_bwd_table_lm one level at a time on the bench workload's surviving samples."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cednerf_b200 as cb
from cednerf_b200 import workload as W, _lib
from cednerf_b200.utils import Rays, _field_fns

cfg, dev = W.DYNERF, torch.device("cuda:0")
est, fld = W.build_scene(cfg, dev, cb)
fld.train(); est.train()
b = {k: v.to(dev) for k, v in W.draw_batch(cfg, 2 ** 18, torch.Generator().manual_seed(1)).items()}
rays = Rays(b["origins"], b["viewdirs"])
with torch.no_grad():
    sigma_fn, _ = _field_fns(fld, rays, b["timestamps"])
    ridx, t0, t1 = est.sampling(rays.origins, rays.viewdirs, sigma_fn=sigma_fn, near_plane=cfg.near_plane,
                                render_step_size=cfg.render_step_size, stratified=True, cone_angle=cfg.cone_angle,
                                alpha_thre=cfg.alpha_thre, jitter=b["jitter"])
n = ridx.numel()
x = rays.origins[ridx] + rays.viewdirs[ridx] * ((t0 + t1) / 2)[:, None]
aabb = fld.aabb
xn = ((x - aabb[:3]) / (aabb[3:] - aabb[:3])).contiguous()
print("samples", n, "per ray", n / 2 ** 18, "xn range", xn.min().item(), xn.max().item())
lv = fld.hash_encoder.levels if hasattr(fld.hash_encoder, "levels") else None
if lv is None:
    for k, v in vars(fld.hash_encoder).items():
        if isinstance(v, _lib.GridLevels):
            lv = v
L = lv.n_levels
total = sum(lv.size[l] for l in range(L))
g = torch.zeros(total, 2, device=dev)
dy = (torch.randn(L, n, 2, device=dev) * 1e-3).half().contiguous()
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def run(levels, dyp):
    rc = lib.cednerf_hashgrid_bwd_table_lm(xn.data_ptr(), 3, n, ctypes.byref(levels), dyp, g.data_ptr(), st)
    assert rc == 0, lib.cednerf_last_error()


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("all levels: %.3f ms" % timeit(lambda: run(lv, dy.data_ptr())))
for l in range(L):
    one = _lib.GridLevels()
    one.n_levels = 1
    one.scale[0], one.res[0], one.size[0], one.offset[0], one.hashed[0] = lv.scale[l], lv.res[l], lv.size[l], lv.offset[l], lv.hashed[l]
    t = timeit(lambda: run(one, dy[l].data_ptr()))
    print(f"level {l:2d} res {lv.res[l]:5d} size {lv.size[l]:8d} hashed {lv.hashed[l]}: {t*1e3:7.1f} us")
