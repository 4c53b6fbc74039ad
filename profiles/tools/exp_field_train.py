"""Phases of the fused TRAINING forward (`field_train_fwd_kernel`) on the visible samples of a DyNeRF-shaped batch, with the
instrumented library (make -C cednerf_b200/csrc debug; CEDNERF_B200_LIB=cednerf_b200/libcednerf_b200_dbg.so), and the time
of every library entry point of one forward + backward.

    python profiles/tools/exp_field_train.py"""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
import cednerf_b200 as cb
from cednerf_b200 import workload as w, _lib
import bench
DEV = torch.device('cuda:0')
cfg = w.DYNERF; rk = w.render_kwargs(cfg)
est, field = w.build_scene(cfg, DEV, cb, seed=42); est.train(); field.train()
b = {k: v.to(DEV) for k, v in w.draw_batch(cfg, 262144, torch.Generator().manual_seed(1000)).items()}
rays = cb.Rays(b['origins'], b['viewdirs'])
def step():
    rgb, acc, depth, n_s, extra = cb.render_image(field, est, rays, render_bkgd=b['color_bkgd'], timestamps=b['timestamps'],
                                                  jitter=b['jitter'], **rk)
    loss = cb.losses.training_loss(rgb, acc, b['pixels'], extra, acc_entropy_loss=True, weight_rgbper=True, use_feat_predict=True)
    for p in field.parameters(): p.grad = None
    (loss * 1024.0).backward()
    return n_s
for _ in range(3): n_s = step()
torch.cuda.synchronize()
lib = _lib.load()
dbg = hasattr(lib, 'cednerf_debug_train_phase_clocks')
buf = (ctypes.c_ulonglong * 12)()
if dbg: lib.cednerf_debug_train_phase_clocks(buf)
with bench.Instrument(cb, _lib) as ins:
    for _ in range(5): step()
    torch.cuda.synchronize()
agg = {}
for name, a, s, e in ins.rec:
    agg.setdefault(name, [0.0, 0]); agg[name][0] += s.elapsed_time(e); agg[name][1] += 1
print('visible samples', int(n_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]): print(f'{k:40s} {v[0] / 5:8.4f} ms per step {v[1] // 5:3d} launches')
if dbg:
    lib.cednerf_debug_train_phase_clocks(buf)
    names = ['sample + exact Frequency row', 'deformation net (+ tile saves)', 'move / normalise / saves', 'hash gathers + blends',
             'exact time embedding + pad + sync', 'density net (+ input / tile saves)', 'sigma, colour input row', 'colour net',
             'rgb out, predictor Frequency row', 'predictor net', 'huber + latent out']
    tiles = buf[11]; tot = sum(buf[i] for i in range(11))
    for i, nm in enumerate(names): print(f'  {nm:36s} {buf[i] / tiles:9.0f} cycles/tile {100.0 * buf[i] / tot:5.1f} %')
    print(f'  total {tot / tiles:.0f} cycles/tile over {tiles} tiles')
