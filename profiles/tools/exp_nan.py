import sys, torch
sys.path.insert(0,'/root/repo')
import cednerf_b200 as cb
from cednerf_b200 import workload as w, ops
DEV='cuda:0'
def run(cfgname, n_rays):
    cfg=getattr(w,cfgname); rk=w.render_kwargs(cfg)
    est, field = w.build_scene(cfg,DEV,cb,seed=42)
    est.train(); field.train()
    b={k:v.to(DEV) for k,v in w.draw_batch(cfg,n_rays,torch.Generator().manual_seed(11)).items()}
    rgb,acc,depth,n_s,extra=cb.render_image(field,est,cb.Rays(b['origins'],b['viewdirs']),render_bkgd=b['color_bkgd'],timestamps=b['timestamps'],jitter=b['jitter'],**rk)
    ex=extra[0]
    l=torch.nn.functional.mse_loss(rgb,b['pixels'])
    if 'latent_losses' in ex: l=l+ex['latent_losses'].mean()
    (l*1024).backward()
    print(cfgname,n_s,'loss',float(l),{k:(bool(torch.isnan(p.grad).any()), float(p.grad.abs().max())) for k,p in field.named_parameters() if p.grad is not None and p.numel()}, 'latent nan', bool(torch.isnan(ex['latent_losses']).any()) if 'latent_losses' in ex else None, flush=True)
for c in sys.argv[1:]:
    run(c,16384)
# dig: which entries, dependence on the sample count
cfg=w.DYNERF; rk=w.render_kwargs(cfg)
est, field = w.build_scene(cfg,DEV,cb,seed=42); est.train(); field.train()
b={k:v.to(DEV) for k,v in w.draw_batch(cfg,16384,torch.Generator().manual_seed(11)).items()}
rays=cb.Rays(b['origins'],b['viewdirs'])
sig,fn=cb.utils._field_fns(field,rays,b['timestamps'])
ridx,t0,t1=est.sampling(b['origins'],b['viewdirs'],sigma_fn=sig,stratified=True,jitter=b['jitter'],**rk)
print('n',t0.numel())
for n in (t0.numel(), t0.numel()-1, t0.numel()-85, 64000, 128*100, 4096):
    for p in field.parameters(): p.grad=None
    r,a,tt=ridx[:n].contiguous(),t0[:n].contiguous(),t1[:n].contiguous()
    rgb,acc,depth,ex=cb.rendering(a,tt,r,16384,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
    l=torch.nn.functional.mse_loss(rgb,b['pixels'])+ex['latent_losses'].mean()
    (l*1024).backward()
    g=field.mlp_feat_prediction.params.grad
    d=field.mlp_feat_prediction.network.desc
    per=[int(torch.isnan(g[d.param_off[l_]:d.param_off[l_]+d.dim_in[l_]*d.dim_out[l_]]).sum()) for l_ in range(d.n_layers)]
    print(n, 'nan per layer', per, 'of', [d.dim_in[l_]*d.dim_out[l_] for l_ in range(d.n_layers)], 'sel false', int((~ex.get('selector', torch.ones(1,dtype=torch.bool,device=DEV))).sum()) if 'selector' in ex else '-')
n=t0.numel()
stash={}
orig=ops.FieldTrainFunction.forward
def fwd(ctx,*a):
    out=orig(ctx,*a); stash['saved']=ctx.to_save[9]; stash['out']=out; return out
ops.FieldTrainFunction.forward=staticmethod(fwd)
for p in field.parameters(): p.grad=None
rgb,acc,depth,ex=cb.rendering(t0,t1,ridx,16384,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
sigma,rgbs,latent,selector,move=stash['out']
print('latent finite', bool(torch.isfinite(latent).all()), 'max', float(latent.max()), 'rows with latent>1e3', torch.nonzero((latent>1e3).any(-1)).flatten().tolist()[:10])
print('sigma finite', bool(torch.isfinite(sigma).all()), 'rgb finite', bool(torch.isfinite(rgbs).all()), 'move finite', bool(torch.isfinite(move).all()))
import ctypes
from cednerf_b200 import _lib
lib=_lib.load()
# saved layout: o4 offset = total - xn - o4...; recompute offsets in python (mirror of saved_layout)
d=field._field_desc()
def layout(n):
    off=0; o={}
    o['h1']=off; off+=(d.f1.n_layers-1)*n*128
    o['o1']=off; off+=n*32
    o['in2']=off; off+=n*d.f2.dim_in[0]*2
    o['h2']=off; off+=(d.f2.n_layers-1)*n*128
    o['o2']=off; off+=n*32
    o['h3']=off; off+=(d.f3.n_layers-1)*n*128
    o['h4']=off; off+=(d.f4.n_layers-1)*n*128
    o['o4']=off; off+=n*64
    o['xn']=off
    return o
lo=layout(n)
sv=stash['saved']
o4=sv[lo['o4']:lo['o4']+n*64].view(torch.float16).view(n,32).float()
in2=sv[lo['in2']:lo['in2']+n*d.f2.dim_in[0]*2].view(torch.float16).view(n,d.f2.dim_in[0]).float()
print('o4 finite', bool(torch.isfinite(o4).all()), 'absmax', float(o4.abs().max()), ' in2 finite', bool(torch.isfinite(in2).all()), float(in2.abs().max()))
bad=torch.nonzero(~torch.isfinite(o4).all(-1)).flatten()
print('rows with non-finite predictor output', bad.tolist()[:20], 'count', bad.numel())
h4=sv[lo['h4']:lo['h4']+n*128].view(torch.float16).view(n,64).float()
print('h4 finite', bool(torch.isfinite(h4).all()), float(h4.abs().max()))
origb=ops.FieldTrainFunction.backward
def bwd(ctx,*g):
    print('incoming finite:', [None if x is None else (bool(torch.isfinite(x).all()), tuple(x.shape)) for x in g])
    dl=g[2]
    print('  d_latent tail rows finite', bool(torch.isfinite(dl[-200:]).all()), 'absmax', float(dl.abs().max()))
    out=origb(ctx,*g)
    print('outgoing finite:', [None if x is None else bool(torch.isfinite(x).all()) for x in out[:5]])
    return out
ops.FieldTrainFunction.backward=staticmethod(bwd)
l=torch.nn.functional.mse_loss(rgb,b['pixels'])+ex['latent_losses'].mean()
(l*1024).backward()
