"""Time of one warm-up occupancy update (4 x 128^3 cells, 8.4 M field queries): fused per-level launch vs closure."""
import sys, torch
sys.path.insert(0,'/root/repo')
import cednerf_b200 as cb
from cednerf_b200 import workload as w, dp
DEV=torch.device('cuda:0')
cfg=w.DYNERF; rk=w.render_kwargs(cfg)
est, field = w.build_scene(cfg,DEV,cb,seed=42); est.train(); field.train()
rng=dp.SharedRng(1,DEV)
fused=cb.utils.FieldOccEval(field, rk['render_step_size'], rng=rng)
plain=lambda x: field.query_density(x, rng.rand(x.shape[0],1))['density']*rk['render_step_size']
for name,fn in (('closure',plain),('FieldOccEval',fused)):
    for step in (0,256):
        for _ in range(2): est._update(step, fn, rng=rng)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        l0=cb._lib.launch_count()
        e0.record()
        for _ in range(5): est._update(step, fn, rng=rng)
        e1.record(); torch.cuda.synchronize()
        print(f"{name:14s} step {step:3d}: {e0.elapsed_time(e1)/5:.3f} ms per update, {(cb._lib.launch_count()-l0)//5} library launches")
import bench
from cednerf_b200 import _lib
with bench.Instrument(cb,_lib) as ins:
    est._update(0, fused, rng=rng); torch.cuda.synchronize()
for name,a,s,e in ins.rec: print(name, round(s.elapsed_time(e),3))
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(4): j=rng.rand(128**3,3); t=rng.rand(128**3,1)
e1.record(); torch.cuda.synchronize(); print('rand for 4 levels', e0.elapsed_time(e1))
e0.record(); m=est.occs.mean(); e1.record(); torch.cuda.synchronize(); print('mean', e0.elapsed_time(e1))
