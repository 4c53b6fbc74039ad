"""The 4-D (xyz + t key-frame) hash encoder at the reference's full size (16 levels, 2^21 entries x 8 fp16 per hashed level,
hash_encoder_inter.py:279-430): forward and table-gradient time per launch against the algorithmic bytes of SURVEY.md 8d
(forward 16 x 8 x 16 B + 16 B + 64 B per sample; backward 64 B dy + 16 B + 16 x 8 x 16 B of fp32 gradient per sample).

    python profiles/tools/exp_hash4d.py [log2 samples]"""
import json, sys, torch
sys.path.insert(0, '/root/repo')
import cednerf_b200 as cb
from cednerf_b200 import workload as w
DEV = torch.device('cuda:0')
n = 2 ** (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
enc = cb.hash_encoder.HashEncoder4D(max_params=2 ** 21, levels=16, base_res=16.0, max_res=2048.0, seed=3).to(DEV)
peak = json.load(open('/root/repo/MEASURED_PEAKS.json'))['hbm_gbs'] if True else 6536.4
g = torch.Generator().manual_seed(5)
# (a) ray-coherent samples: packed samples of a DyNeRF-shaped training batch, normalised to the unit cube
cfg = w.DYNERF; rk = w.render_kwargs(cfg)
est, field = w.build_scene(cfg, DEV, cb, seed=42); est.train()
b = {k: v.to(DEV) for k, v in w.draw_batch(cfg, 262144, torch.Generator().manual_seed(1000)).items()}
with torch.no_grad():
    ridx, t0, t1 = est.sampling(b['origins'], b['viewdirs'], stratified=True, jitter=b['jitter'], **rk)
x = b['origins'][ridx] + b['viewdirs'][ridx] * ((t0 + t1) / 2)[:, None]
lo, hi = torch.tensor(cfg.roi_aabb[:3], device=DEV), torch.tensor(cfg.roi_aabb[3:], device=DEV)
coherent = torch.cat([((x - lo) / (hi - lo)).clamp(0, 1), b['timestamps'][ridx]], -1)[:n].contiguous()
uniform = torch.rand(coherent.shape[0], 4, generator=g).to(DEV)
del est, field
def timed(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for name, pts in (('ray-coherent', coherent), ('uniform', uniform)):
    m = pts.shape[0]
    dy = torch.randn(m, 32, device=DEV).half()
    with torch.no_grad():
        fwd = timed(lambda: enc(pts))
    y = enc(pts)
    def bwd():
        enc.hash_table.grad = None
        y.backward(dy, retain_graph=True)
    tot = timed(bwd, 5)
    zero = timed(lambda: torch.zeros_like(enc.hash_table), 5)
    fb, bb = 2128.0 * m, (64.0 + 16.0 + 2048.0) * m
    print(f'{name:13s} {m} samples: forward {fwd:.3f} ms = {fb / fwd / 1e6:.0f} GB/s ({fb / fwd / 1e6 / peak:.2f} of {peak:.0f}); '
          f'backward {tot:.3f} ms incl. {zero:.3f} ms zero fill of the 766 MB fp32 gradient -> kernel {tot - zero:.3f} ms = '
          f'{bb / (tot - zero) / 1e6:.0f} GB/s ({bb / (tot - zero) / 1e6 / peak:.2f})')
