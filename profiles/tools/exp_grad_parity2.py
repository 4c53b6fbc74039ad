"""Where does the deformation-net gradient of the fused training path leave the oracle?  Compares, on the oracle's sample
set: oracle (CPU) vs the op-by-op CUDA path vs the fused CUDA path."""
import sys, torch
sys.path.insert(0,'/root/repo')
from oracle import cednerf_ref as cr, nerfacc_ref as nf
import cednerf_b200 as cb
from cednerf_b200 import workload as w
DEV='cuda:0'
class _O: OccGridEstimator, DNGPradianceField = nf.OccGridEstimator, cr.DNGPradianceField
def rel(a,b): return float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
cfgname=sys.argv[1]; n_rays=int(sys.argv[2])
cfg=getattr(w,cfgname); rk=w.render_kwargs(cfg)
est_ref, ref = w.build_scene(cfg,'cpu',_O,seed=42)
est, field = w.build_scene(cfg,DEV,cb,seed=42)
for m in (est,field,est_ref,ref): m.train()
batch=w.draw_batch(cfg,n_rays,torch.Generator().manual_seed(11)); b={k:v.to(DEV) for k,v in batch.items()}
scale=float(sys.argv[3]) if len(sys.argv)>3 else 1024.0
out_ref=cr.render_image(ref,est_ref,cr.Rays(batch['origins'],batch['viewdirs']),render_bkgd=batch['color_bkgd'],timestamps=batch['timestamps'],jitter=batch['jitter'],**rk)
ex_ref=out_ref[4][0]
def loss_of(r,e,p):
    l=torch.nn.functional.mse_loss(r,p)
    if 'latent_losses' in e: l=l+e['latent_losses'].mean()
    return l
(loss_of(out_ref[0],ex_ref,batch['pixels'])*scale).backward()
gref={k:q.grad.clone() for k,q in ref.named_parameters() if q.grad is not None and q.numel()}
t0,t1,ridx=ex_ref['t_starts'].to(DEV),ex_ref['t_ends'].to(DEV),ex_ref['ray_indices'].to(DEV)
res={}
for mode in ('fused','opbyop'):
    for p in field.parameters(): p.grad=None
    if mode=='opbyop': field.fused_train_supported=lambda: False
    _,fn=cb.utils._field_fns(field,cb.Rays(b['origins'],b['viewdirs']),b['timestamps'])
    rgb,acc,depth,ex=cb.rendering(t0,t1,ridx,n_rays,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
    (loss_of(rgb,ex,b['pixels'])*scale).backward()
    res[mode]={k:p.grad.detach().cpu().clone() for k,p in field.named_parameters() if p.grad is not None and p.numel()}
    print(mode,{k:f"{rel(res[mode][k],gref[k]):.2e}" for k in gref}, flush=True)
print('fused vs opbyop',{k:f"{rel(res['fused'][k],res['opbyop'][k]):.2e}" for k in gref})
# per layer of the deformation net
d=field.xyz_wrap.network.desc
for l in range(d.n_layers):
    o,n=d.param_off[l], d.dim_in[l]*d.dim_out[l]
    print('layer',l,'fused',f"{rel(res['fused']['xyz_wrap.params'][o:o+n],gref['xyz_wrap.params'][o:o+n]):.2e}",'opbyop',f"{rel(res['opbyop']['xyz_wrap.params'][o:o+n],gref['xyz_wrap.params'][o:o+n]):.2e}")
# determinism of the fused backward (shared TMEM weight-gradient accumulators): two runs on identical inputs
runs=[]
field.fused_train_supported=type(field).fused_train_supported.__get__(field)
for _ in range(2):
    for p in field.parameters(): p.grad=None
    _,fn=cb.utils._field_fns(field,cb.Rays(b['origins'],b['viewdirs']),b['timestamps'])
    rgb,acc,depth,ex=cb.rendering(t0,t1,ridx,n_rays,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
    (loss_of(rgb,ex,b['pixels'])*scale).backward()
    runs.append({k:p.grad.detach().cpu().clone() for k,p in field.named_parameters() if p.grad is not None and p.numel()})
print('run-to-run',{k:f"{rel(runs[0][k],runs[1][k]):.2e}" for k in gref})
