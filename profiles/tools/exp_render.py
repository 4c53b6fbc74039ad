"""Per-round timeline of one 1352x1014 frame through render_image_test (device-resident rounds): kernel time by entry
point (march_round split into count / fallback fill), alive rays and k per round."""
import sys, torch
sys.path.insert(0,'/root/repo')
import cednerf_b200 as cb
from cednerf_b200 import workload as w, _lib, utils as U
import bench
DEV=torch.device('cuda:0')
cfg=w.DYNERF; rk=w.render_kwargs(cfg)
est, field = w.build_scene(cfg,DEV,cb,seed=42); est.eval(); field.eval()
pose=w.spiral_poses(cfg,300)[37]
o,d=w.pose_rays(cfg,pose,False,DEV)
rays=cb.Rays(o,d); t=torch.tensor([[0.123]],device=DEV); bk=torch.zeros(3,device=DEV)
for _ in range(2): cb.render_image_test(1024,field,est,rays,render_bkgd=bk,timestamps=t,**rk)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): n_s=cb.render_image_test(1024,field,est,rays,render_bkgd=bk,timestamps=t,**rk)[3]
e1.record(); torch.cuda.synchronize()
print('ms/frame',e0.elapsed_time(e1)/5,'samples',n_s)
with bench.Instrument(cb,_lib) as ins:
    cb.render_image_test(1024,field,est,rays,render_bkgd=bk,timestamps=t,**rk); torch.cuda.synchronize()
agg={}
rounds=[]
for name,a,s,e in ins.rec:
    key=name+('.fill' if name=='cednerf_march_round' and a[0]==1 else '')
    agg.setdefault(key,[0.0,0]); agg[key][0]+=s.elapsed_time(e); agg[key][1]+=1
    if name=='cednerf_march_round' and a[0]==0: rounds.append((a[3], round(s.elapsed_time(e),3)))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0]): print(f"{k:40s} {v[0]:8.3f} ms {v[1]:4d}")
print('count pass per round (bound, ms):', rounds)
U._DEVICE_ROUNDS=False
ref=cb.render_image_test(1024,field,est,rays,render_bkgd=bk,timestamps=t,**rk)
U._DEVICE_ROUNDS=True
got=[cb.render_image_test(1024,field,est,rays,render_bkgd=bk,timestamps=t,**rk) for _ in range(3)]
print('host-driven total',ref[3],'device rounds totals',[g[3] for g in got],'max |rgb diff|',[float((g[0]-ref[0]).abs().max()) for g in got], 'opacity diff', float((got[0][1]-ref[1]).abs().max()))
