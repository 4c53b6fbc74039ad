"""Bisect the deformation-net gradient of the fused backward: saved activations vs the op-by-op path's, and weight
gradients re-derived in fp64 from the op-by-op path's masked hidden gradients."""
import sys, torch
sys.path.insert(0,'/root/repo')
from oracle import cednerf_ref as cr, nerfacc_ref as nf
import cednerf_b200 as cb
from cednerf_b200 import workload as w, ops
DEV='cuda:0'
class _O: OccGridEstimator, DNGPradianceField = nf.OccGridEstimator, cr.DNGPradianceField
def rel(a,b): return float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
cfgname=sys.argv[1]; n_rays=int(sys.argv[2])
cfg=getattr(w,cfgname); rk=w.render_kwargs(cfg)
est_ref, ref = w.build_scene(cfg,'cpu',_O,seed=42)
est, field = w.build_scene(cfg,DEV,cb,seed=42)
for m in (est,field,est_ref,ref): m.train()
batch=w.draw_batch(cfg,n_rays,torch.Generator().manual_seed(11)); b={k:v.to(DEV) for k,v in batch.items()}
scale=1024.0
with torch.no_grad():
    sig_fn,_=cr._field_fns(ref, cr.Rays(batch['origins'],batch['viewdirs']), batch['timestamps'])
ridx,t0,t1=est_ref.sampling(batch['origins'],batch['viewdirs'],sigma_fn=sig_fn,stratified=True,jitter=batch['jitter'],**rk)
t0,t1,ridx=t0.to(DEV),t1.to(DEV),ridx.to(DEV)
n=t0.numel(); print('samples',n)
def loss_of(r,e,p):
    l=torch.nn.functional.mse_loss(r,p)
    if 'latent_losses' in e: l=l+e['latent_losses'].mean()
    return l
# fused, stashing `saved`
stash={}
orig_fwd=ops.FieldTrainFunction.forward
def fwd(ctx,*a):
    out=orig_fwd(ctx,*a); stash['saved']=ctx.to_save[9]; return out
ops.FieldTrainFunction.forward=staticmethod(fwd)
_,fn=cb.utils._field_fns(field,cb.Rays(b['origins'],b['viewdirs']),b['timestamps'])
rgb,acc,depth,ex=cb.rendering(t0,t1,ridx,n_rays,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
(loss_of(rgb,ex,b['pixels'])*scale).backward()
g_fused=field.xyz_wrap.params.grad.detach().clone()
h_f=stash['saved'][:3*n*128].view(torch.float16).view(3,n,64)
# op-by-op with debug capture
for p in field.parameters(): p.grad=None
field.fused_train_supported=lambda: False
ops.MlpFunction._debug={}
caps={}
orig_mf=ops.MlpFunction.forward
def mf(ctx,x,params,image,desc,save,n_out):
    out=orig_mf(ctx,x,params,image,desc,save,n_out)
    if save and desc.n_layers==4: caps['x16'],caps['hidden']=ctx.to_save[0],ctx.to_save[1]
    return out
ops.MlpFunction.forward=staticmethod(mf)
_,fn=cb.utils._field_fns(field,cb.Rays(b['origins'],b['viewdirs']),b['timestamps'])
rgb,acc,depth,ex=cb.rendering(t0,t1,ridx,n_rays,rgb_sigma_fn=fn,render_bkgd=b['color_bkgd'])
(loss_of(rgb,ex,b['pixels'])*scale).backward()
g_op=field.xyz_wrap.params.grad.detach().clone()
dh=[t for t in ops.MlpFunction._debug['d_hidden'] if t.shape[0]==3][-1]   # deformation net: [3, n, 64] masked hidden grads
h_o=caps['hidden']
print('saved activations fused vs op-by-op: equal', [bool(torch.equal(h_f[l],h_o[l])) for l in range(3)],
      [float((h_f[l].float()-h_o[l].float()).abs().max()) for l in range(3)])
d=field.xyz_wrap.network.desc
ins=[caps['x16'],h_o[0],h_o[1]]
for l in range(3):
    o,k,m=d.param_off[l],d.dim_in[l],d.dim_out[l]
    want=(dh[l].double().T@ins[l].double()).float().reshape(-1)
    print('layer',l,'fp64 re-derivation vs op-by-op',f"{rel(g_op[o:o+m*k],want):.2e}",'vs fused',f"{rel(g_fused[o:o+m*k],want):.2e}",
          ' |dh| max',float(dh[l].abs().max()),'nonzero frac',float((dh[l]!=0).float().mean()), 'subnormal frac', float(((dh[l]!=0)&(dh[l].abs()<6.1e-5)).float().mean()))
# structure of the difference
for l in range(3):
    o,k,m=d.param_off[l],d.dim_in[l],d.dim_out[l]
    a_,b_=g_fused[o:o+m*k].double().view(m,k), g_op[o:o+m*k].double().view(m,k)
    diff=a_-b_
    alpha=float((a_*b_).sum()/(b_*b_).sum())
    print('layer',l,'|op|',float(b_.norm()),'scale fit',alpha,'residual after scale',float((a_-alpha*b_).norm()/b_.norm()),
          'row-norm of diff (first 8 out-neurons)',[f"{float(x):.1e}" for x in diff.norm(dim=1)[:8]],
          'col-norm (first 8 inputs)',[f"{float(x):.1e}" for x in diff.norm(dim=0)[:8]])
    rn=diff.norm(dim=1); cn=diff.norm(dim=0)
    print('   rows with largest diff',torch.topk(rn,4).indices.tolist(),[f"{float(x):.1e}" for x in torch.topk(rn,4).values],' cols',torch.topk(cn,4).indices.tolist(),[f"{float(x):.1e}" for x in torch.topk(cn,4).values])
