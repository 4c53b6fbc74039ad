"""The sigma pre-pass of one DyNeRF-shaped training batch (2^18 rays, ~9.1 M marched samples) on its own: time per launch
and, with the instrumented library (make -C cednerf_b200/csrc debug; CEDNERF_B200_LIB=cednerf_b200/libcednerf_b200_dbg.so),
the cycles a 128-sample tile spends in each phase of the fused kernel.

    python profiles/tools/exp_field_fwd.py [n_rays]"""
import ctypes, os, sys, torch
sys.path.insert(0, '/root/repo')
import cednerf_b200 as cb
from cednerf_b200 import workload as w, ops, _lib
DEV = torch.device('cuda:0')
cfg = w.DYNERF; rk = w.render_kwargs(cfg)
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
est, field = w.build_scene(cfg, DEV, cb, seed=42); est.train(); field.train()
b = {k: v.to(DEV) for k, v in w.draw_batch(cfg, n_rays, torch.Generator().manual_seed(1000)).items()}
captured = {}
orig = ops.field_fwd
def spy(*a, **kw):
    if a[3] > captured.get('n', 0):
        captured.update(n=a[3], args=a, kw=kw)
    return orig(*a, **kw)
ops.field_fwd = spy
with torch.no_grad():
    sigma_fn = cb.utils._field_fns(field, cb.Rays(b['origins'], b['viewdirs']), b['timestamps'])[0]
    ridx, t0, t1 = est.sampling(b['origins'], b['viewdirs'], sigma_fn=sigma_fn, stratified=True, jitter=b['jitter'], **rk)
ops.field_fwd = orig
torch.cuda.synchronize()
n = captured['n']
print('marched samples', n, 'visible', ridx.numel())
def run(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): out = orig(*captured['args'], **captured['kw'])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, out
run(3)
lib = _lib.load() if hasattr(_lib, 'load') else None
dbg = lib is not None and hasattr(lib, 'cednerf_debug_phase_clocks')
buf = (ctypes.c_ulonglong * 8)()
if dbg: lib.cednerf_debug_phase_clocks(buf)
ms, out = run(10)
print(f'field_fwd: {ms:.4f} ms per launch, {n / ms / 1e3:.1f} M samples/s, sigma checksum {float(out[0][:n].double().sum()):.6e}')
if dbg:
    lib.cednerf_debug_phase_clocks(buf)
    tiles = buf[7]
    names = ['sample+frequency', 'deformation net', 'move/normalise', 'hash gathers', 'time emb+pad+sync', 'density net', 'sigma out']
    tot = sum(buf[i] for i in range(7))
    for i, nm in enumerate(names):
        print(f'  {nm:20s} {buf[i] / tiles:9.0f} cycles/tile  {100.0 * buf[i] / tot:5.1f} %')
    print(f'  total {tot / tiles:.0f} cycles/tile over {tiles} tiles')
