#!/usr/bin/env python
"""Diagnostic (GPU): cost of OccGridEstimator._update on the bench scene (warm-up step: all cells; later: R^3/4 + occupied)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cednerf_b200 as cb
from cednerf_b200 import workload as W, _lib

cfg, dev = W.DYNERF, torch.device("cuda:0")
est, fld = W.build_scene(cfg, dev, cb)
fld.train(); est.train()


def occ_eval_fn(x):  # train_real.py:324-328
    t = torch.rand(x.shape[0], 1, device=x.device)
    with torch.no_grad():
        return fld.query_density(x, t)["density"] * cfg.render_step_size


for step in (0, 1024):
    for rep in range(3):
        torch.cuda.synchronize(); l0 = _lib.launch_count(); t0 = time.perf_counter()
        est._update(step, occ_eval_fn)
        torch.cuda.synchronize()
        print(f"step {step}: {(time.perf_counter() - t0) * 1e3:.2f} ms, {_lib.launch_count() - l0} library launches")
if len(sys.argv) > 1:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        est._update(1024, occ_eval_fn); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
