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
for scale in (1024.0, 65536.0):
    for p in list(ref.parameters())+list(field.parameters()): p.grad=None
    out_ref=cr.render_image(ref,est_ref,cr.Rays(batch['origins'],batch['viewdirs']),render_bkgd=batch['color_bkgd'],timestamps=batch['timestamps'],jitter=batch['jitter'],**rk)
    ex_ref=out_ref[4][0]
    _,fn=cb.utils._field_fns(field,cb.Rays(b['origins'],b['viewdirs']),b['timestamps'])
    t0,t1,ridx=ex_ref['t_starts'].to(DEV),ex_ref['t_ends'].to(DEV),ex_ref['ray_indices'].to(DEV)
    rgb,acc,depth,ex=cb.rendering(t0,t1,ridx,n_rays,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
    def loss_of(r,e,p):
        l=torch.nn.functional.mse_loss(r,p)
        if 'latent_losses' in e: l=l+e['latent_losses'].mean()
        return l
    (loss_of(out_ref[0],ex_ref,batch['pixels'])*scale).backward(); (loss_of(rgb,ex,b['pixels'])*scale).backward()
    print(cfgname,n_rays,scale,out_ref[3],{k:f"{rel(p.grad.cpu(),q.grad):.2e}" for (k,p),(_,q) in zip(field.named_parameters(),ref.named_parameters()) if q.grad is not None and q.numel()}, flush=True)
    g=ref.xyz_wrap.params.grad; print('  |g_xyz| ref', float(g.norm()), 'max', float(g.abs().max()))
