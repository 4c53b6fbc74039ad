"""GPU parity: every CUDA kernel family against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): bit-exact for integer / index outputs (per-ray sample counts, packed
ray_indices, interval flags) and for the packed fp32 t values the marcher emits; fp32 compositing
max-abs <= 1e-4; fp16 MLP path <= 2e-3; gradients relative error <= 1e-3.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import nerfacc_ref as nf  # noqa: E402
from oracle import taichi_ref as tr  # noqa: E402
from oracle import tcnn_ref as tc  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def cb():
    import cednerf_b200

    return cednerf_b200


def make_scene(seed, n_rays, res, levels, occ=0.15, inside=False):
    g = torch.Generator().manual_seed(seed)
    est = nf.OccGridEstimator([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], resolution=res, levels=levels)
    est.binaries = torch.rand(levels, res, res, res, generator=g) < occ
    est.occs = est.binaries.flatten().float() * 0.5
    if inside:
        origins = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.5
    else:
        origins = torch.tensor([0.0, 0.0, -3.5]) + (torch.rand(n_rays, 3, generator=g) - 0.5) * 0.6
    target = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.6
    dirs = torch.nn.functional.normalize(target - origins, dim=-1)
    return est, origins, dirs, g


# ---------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("levels,res,step,cone,inside", [(1, 16, 2e-2, 0.0, False), (2, 16, 2e-2, 0.004, False),
                                                          (4, 32, 1e-2, 0.004, True), (4, 128, 5e-3, 0.004, False)])
def test_march_two_pass_bit_exact(cb, levels, res, step, cone, inside):
    est, o, d, g = make_scene(3 + levels, 2048, res, levels, occ=0.1 if res < 128 else 0.02, inside=inside)
    near = torch.full((o.shape[0],), 0.2) + torch.rand(o.shape[0], generator=g) * step
    far = torch.full((o.shape[0],), 1e10)
    ridx, t0, t1, packed, term = nf.traverse_grids(o, d, est.binaries, est.aabbs, near, far, step, cone,
                                                   packed_only=True)
    assert ridx.numel() > 1000
    ops = cb.ops
    bits = ops.pack_occupancy(est.binaries.to(DEV))
    mi = ops.MarchInputs(o.to(DEV), d.to(DEV), bits, est.aabbs.to(DEV), res, near.to(DEV), far.to(DEV), 0.0, 1e10,
                         step, cone)
    n_iv, n_sm, term_g = mi.count()
    starts, packed_g, total = ops.exclusive_scan(n_sm)
    assert int(total.item()) == ridx.numel()
    assert torch.equal(packed_g.cpu(), packed)
    ridx_g, t0_g, t1_g, term_f = mi.fill_packed(starts, int(total.item()))
    assert torch.equal(ridx_g.cpu(), ridx)
    assert torch.equal(t0_g.cpu(), t0) and torch.equal(t1_g.cpu(), t1)
    assert torch.equal(term_g.cpu(), term) and torch.equal(term_f.cpu(), term)
    # properties: sorted by ray then t, intervals non-empty
    assert bool((ridx_g[1:] >= ridx_g[:-1]).all()) and bool((t1_g > t0_g).all())


@pytest.mark.parametrize("cap", [1, 3, 16])
def test_march_fill_from_recorded_runs(cb, cap, monkeypatch):
    """The fill pass that replays the count pass's runs (and its full-march fallback for rays with more than `cap`
    runs) emits exactly the oracle's packed samples."""
    est, o, d, g = make_scene(17, 3000, 16, 2, occ=0.25)
    near = torch.full((o.shape[0],), 0.2) + torch.rand(o.shape[0], generator=g) * 2e-2
    far = torch.full((o.shape[0],), 1e10)
    ridx, t0, t1, packed, _ = nf.traverse_grids(o, d, est.binaries, est.aabbs, near, far, 2e-2, 0.004, packed_only=True)
    ops = cb.ops
    monkeypatch.setattr(ops.MarchInputs, "RUN_CAP", cap)
    mi = ops.MarchInputs(o.to(DEV), d.to(DEV), ops.pack_occupancy(est.binaries.to(DEV)), est.aabbs.to(DEV), 16,
                         near.to(DEV), far.to(DEV), 0.0, 1e10, 2e-2, 0.004)
    _, n_sm, _ = mi.count(record_runs=True)
    n_runs = mi.runs[2]
    assert int(n_runs.max()) > cap or cap == 16            # the fallback is exercised for the small capacities
    starts, packed_g, total = ops.exclusive_scan(n_sm)
    r_g, a_g, b_g = mi.fill_packed_from_runs(starts, int(total.item()))
    assert torch.equal(packed_g.cpu(), packed)
    assert torch.equal(r_g.cpu(), ridx) and torch.equal(a_g.cpu(), t0) and torch.equal(b_g.cpu(), t1)


def test_ray_aabb_and_sort(cb):
    est, o, d, _ = make_scene(11, 4096, 16, 4)
    o[:7] = 0.0  # origins inside all boxes
    d[5] = torch.tensor([0.0, 0.0, 1.0])  # axis-aligned ray (division by zero in the slab test)
    t_mins, t_maxs, hits = nf.ray_aabb_intersect(o, d, est.aabbs)
    ts, ti = nf.sort_boundaries(t_mins, t_maxs)
    a, b, h = cb.nerfacc.ray_aabb_intersect(o.to(DEV), d.to(DEV), est.aabbs.to(DEV))
    assert torch.equal(a.cpu(), t_mins) and torch.equal(b.cpu(), t_maxs) and torch.equal(h.cpu(), hits)
    ts_g, ti_g = cb.ops.sort_boundaries(a, b)
    assert torch.equal(ts_g.cpu(), ts) and torch.equal(ti_g.cpu(), ti)


@pytest.mark.parametrize("limit,cone", [(4, 0.004), (1, 0.0), (64, 0.0)])
def test_traverse_grids_limited_over_allocate(cb, limit, cone):
    est, o, d, g = make_scene(21, 1024, 16, 2, occ=0.2)
    n = o.shape[0]
    near, far = torch.full((n,), 0.2), torch.full((n,), 1e10)
    mask = torch.rand(n, generator=g) < 0.7
    t_mins, t_maxs, hits = nf.ray_aabb_intersect(o, d, est.aabbs)
    ts, ti = nf.sort_boundaries(t_mins, t_maxs)
    iv, sm, term = nf.traverse_grids(o, d, est.binaries, est.aabbs, near, far, 2e-2, cone, limit, True, mask, ts, ti,
                                     hits)
    to = lambda t: t.to(DEV)
    iv_g, sm_g, term_g = cb.nerfacc.traverse_grids(to(o), to(d), to(est.binaries), to(est.aabbs), to(near), to(far),
                                                   2e-2, cone, limit, True, to(mask), to(ts), to(ti), to(hits))
    for a, b in ((iv_g.vals, iv.vals), (iv_g.is_left, iv.is_left), (iv_g.is_right, iv.is_right),
                 (iv_g.packed_info, iv.packed_info), (sm_g.vals, sm.vals), (sm_g.ray_indices, sm.ray_indices),
                 (sm_g.is_valid, sm.is_valid), (sm_g.packed_info, sm.packed_info)):
        assert torch.equal(a.cpu(), b)
    assert torch.equal(term_g.cpu()[mask], term[mask])
    # second round from the termination planes (the reference's render_image_test loop)
    iv2, sm2, _ = nf.traverse_grids(o, d, est.binaries, est.aabbs, term, far, 2e-2, cone, limit, True, mask, ts, ti, hits)
    iv2_g, sm2_g, _ = cb.nerfacc.traverse_grids(to(o), to(d), to(est.binaries), to(est.aabbs), term_g, to(far), 2e-2,
                                                cone, limit, True, to(mask), to(ts), to(ti), to(hits))
    assert torch.equal(iv2_g.vals.cpu(), iv2.vals) and torch.equal(sm2_g.is_valid.cpu(), sm2.is_valid)


def test_traverse_grids_two_pass_nerfacc_shape_and_empty(cb):
    est, o, d, _ = make_scene(5, 512, 16, 2)
    to = lambda t: t.to(DEV)
    iv, sm, term = nf.traverse_grids(o, d, est.binaries, est.aabbs, None, None, 2e-2, 0.0)
    iv_g, sm_g, term_g = cb.nerfacc.traverse_grids(to(o), to(d), to(est.binaries), to(est.aabbs), None, None, 2e-2, 0.0)
    for a, b in ((iv_g.vals, iv.vals), (iv_g.is_left, iv.is_left), (iv_g.is_right, iv.is_right),
                 (iv_g.ray_indices, iv.ray_indices), (iv_g.packed_info, iv.packed_info), (sm_g.vals, sm.vals),
                 (sm_g.ray_indices, sm.ray_indices), (sm_g.packed_info, sm.packed_info), (term_g, term)):
        assert torch.equal(a.cpu(), b)
    # empty grid, rays that miss everything, zero rays
    est.binaries = torch.zeros_like(est.binaries)
    iv_g, sm_g, _ = cb.nerfacc.traverse_grids(to(o), to(d), to(est.binaries), to(est.aabbs), None, None, 2e-2, 0.0)
    assert iv_g.vals.numel() == 0 and sm_g.vals.numel() == 0 and int(sm_g.packed_info[:, 1].sum()) == 0
    iv_g, sm_g, t_g = cb.nerfacc.traverse_grids(to(o[:0]), to(d[:0]), to(est.binaries), to(est.aabbs), None, None, 2e-2, 0.0)
    assert sm_g.vals.numel() == 0 and t_g.numel() == 0


def test_occ_threshold_and_pack(cb):
    g = torch.Generator().manual_seed(1)
    occs = torch.rand(2 * 32 ** 3, generator=g)
    thre = torch.tensor([0.37])
    bins = torch.empty(2, 32, 32, 32, dtype=torch.bool, device=DEV)
    bits = torch.empty(occs.numel() // 32, dtype=torch.int32, device=DEV)
    cb.ops.occ_threshold_pack(occs.to(DEV), thre.to(DEV), bins, bits)
    want = (occs > 0.37).view(2, 32, 32, 32)
    assert torch.equal(bins.cpu(), want)
    assert torch.equal(cb.ops.pack_occupancy(want.to(DEV)).cpu(), bits.cpu())
    w = want.flatten().view(-1, 32).to(torch.int64)
    ref_words = (w << torch.arange(32)).sum(-1)
    assert torch.equal(bits.cpu().to(torch.int64) & 0xFFFFFFFF, ref_words)


# ---------------------------------------------------------------------------------------------- K4
def _samples(seed, n_rays, mean, empty_frac=0.3):
    g = torch.Generator().manual_seed(seed)
    cnt = torch.poisson(torch.full((n_rays,), float(mean)), generator=g).long()
    cnt[torch.rand(n_rays, generator=g) < empty_frac] = 0
    ridx = torch.repeat_interleave(torch.arange(n_rays), cnt)
    s = ridx.numel()
    dt = torch.rand(s, generator=g) * 0.05 + 1e-3
    t0 = torch.rand(s, generator=g) * 3
    sig = torch.rand(s, generator=g) * 40 * (torch.rand(s, generator=g) < 0.7)
    rgb = torch.rand(s, 3, generator=g)
    return g, ridx, t0, t0 + dt, sig, rgb


@pytest.mark.parametrize("mean", [3, 10, 20, 90])
def test_composite_fused_fwd_bwd(cb, mean):
    n_rays = 777
    g, ridx, t0, t1, sig, rgb = _samples(mean, n_rays, mean)
    bk = torch.rand(3, generator=g)
    gc, go, gd = torch.rand(n_rays, 3, generator=g), torch.rand(n_rays, 1, generator=g), torch.rand(n_rays, 1, generator=g)
    gw, gt = torch.rand(ridx.numel(), generator=g), torch.rand(ridx.numel(), generator=g)

    def run(dev, fn):
        dt = torch.float64 if dev == "cpu" else torch.float32
        s_ = sig.to(dev, dt).requires_grad_(True)
        r_ = rgb.to(dev, dt).requires_grad_(True)
        outs = fn(s_, r_)
        loss = sum((o * w.to(dev)).sum() for o, w in zip(outs[:5], (gc, go, gd, gw, gt)))
        loss.backward()
        return [o.detach().float().cpu() for o in outs], s_.grad.float().cpu(), r_.grad.float().cpu()

    def oracle(s_, r_):
        w, tr_, al = nf.render_weight_from_density(t0.double(), t1.double(), s_, ray_indices=ridx, n_rays=n_rays)
        c = nf.accumulate_along_rays(w, r_, ridx, n_rays)
        o = nf.accumulate_along_rays(w, None, ridx, n_rays)
        dp = nf.accumulate_along_rays(w, ((t0 + t1) / 2).double()[:, None], ridx, n_rays)
        dp = dp / o.clamp_min(torch.finfo(torch.float32).eps)
        return c + bk.double() * (1 - o), o, dp, w, tr_, al

    def ours(s_, r_):
        off = cb.ops.ray_offsets(ridx.to(DEV), n_rays)
        return cb.ops.CompositeFunction.apply(t0.to(DEV), t1.to(DEV), s_, r_, off, n_rays, bk.to(DEV))

    o_ref, gs_ref, gr_ref = run("cpu", oracle)
    o_gpu, gs_gpu, gr_gpu = run(DEV, ours)
    for a, b in zip(o_gpu, o_ref):
        torch.testing.assert_close(a, b, rtol=0, atol=1e-4)  # north-star: fp32 rgb/opacity/depth max-abs <= 1e-4
    torch.testing.assert_close(gr_gpu, gr_ref, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(gs_gpu, gs_ref, rtol=1e-3, atol=1e-5)


def test_volrend_api_and_accumulate(cb):
    n_rays = 300
    g, ridx, t0, t1, sig, rgb = _samples(99, n_rays, 12)
    to = lambda t: t.to(DEV)
    prefix = torch.rand(ridx.numel(), generator=g)
    w, tr_, al = nf.render_weight_from_density(t0, t1, sig, ray_indices=ridx, n_rays=n_rays, prefix_trans=prefix)
    w_g, tr_g, al_g = cb.nerfacc.render_weight_from_density(to(t0), to(t1), to(sig), ray_indices=to(ridx),
                                                           n_rays=n_rays, prefix_trans=to(prefix))
    for a, b in ((w_g, w), (tr_g, tr_), (al_g, al)):
        torch.testing.assert_close(a.cpu(), b, rtol=0, atol=1e-5)
    packed = nf.packed_info_from_indices(ridx, n_rays)
    tr2, al2 = cb.nerfacc.render_transmittance_from_density(to(t0), to(t1), to(sig), packed_info=to(packed))
    torch.testing.assert_close(tr2.cpu(), nf.render_transmittance_from_density(t0, t1, sig, ray_indices=ridx)[0],
                               rtol=0, atol=1e-5)
    vis = nf.render_visibility_from_density(t0, t1, sig, ray_indices=ridx, n_rays=n_rays, early_stop_eps=1e-2,
                                            alpha_thre=0.05)
    vis_g = cb.nerfacc.render_visibility_from_density(to(t0), to(t1), to(sig), ray_indices=to(ridx), n_rays=n_rays,
                                                      early_stop_eps=1e-2, alpha_thre=0.05)
    assert (vis_g.cpu() != vis).float().mean() < 1e-3  # threshold ties only
    # accumulate: out-of-place with autograd, and in place
    for n_ch in (5, 12, 32):  # narrow (group-per-ray) and wide (lane = channel) kernels
        vals = torch.rand(ridx.numel(), n_ch, generator=g)
        gout = torch.rand(n_rays, n_ch, generator=g)
        wv, vv = w.clone().requires_grad_(True), vals.clone().requires_grad_(True)
        (nf.accumulate_along_rays(wv, vv, ridx, n_rays) * gout).sum().backward()
        wg, vg = to(w).requires_grad_(True), to(vals).requires_grad_(True)
        out_g = cb.nerfacc.accumulate_along_rays(wg, vg, to(ridx), n_rays)
        (out_g * to(gout)).sum().backward()
        torch.testing.assert_close(out_g.detach().cpu(), nf.accumulate_along_rays(w, vals, ridx, n_rays), rtol=1e-5,
                                   atol=1e-5)
        torch.testing.assert_close(wg.grad.cpu(), wv.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(vg.grad.cpu(), vv.grad, rtol=1e-5, atol=1e-6)
    acc = torch.rand(n_rays, 1, generator=g)
    acc_g = to(acc).clone()
    cb.nerfacc.accumulate_along_rays_(to(w), None, to(ridx), acc_g)
    nf.accumulate_along_rays_(w, None, ridx, acc)
    torch.testing.assert_close(acc_g.cpu(), acc, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------- K2
GRID_CASES = {"small_hashed": dict(n_levels=8, base=16, dst=256, log2_t=12),
              "dnerf": dict(n_levels=16, base=16, dst=1024, log2_t=19),
              "dynerf": dict(n_levels=16, base=16, dst=8192, log2_t=21)}


@pytest.mark.parametrize("case", list(GRID_CASES))
def test_hashgrid_fwd_bwd(cb, case):
    c = GRID_CASES[case]
    b = math.exp(math.log(c["dst"] / c["base"]) / (c["n_levels"] - 1))
    cfg = {"otype": "HashGrid", "n_levels": c["n_levels"], "n_features_per_level": 2, "log2_hashmap_size": c["log2_t"],
           "base_resolution": c["base"], "per_level_scale": b}
    ref = tc.Encoding(3, cfg, seed=7)
    enc = cb.tcnn.Encoding(3, cfg, seed=7).to(DEV)
    lv = ref.levels
    assert [i[1] for i in enc.level_info] == lv[1] and [i[2] for i in enc.level_info] == lv[2]
    assert [i[3] for i in enc.level_info] == lv[3] and [bool(i[4]) for i in enc.level_info] == lv[4]
    if case == "dynerf":
        assert lv[1][:6] == [16, 25, 37, 56, 85, 128] and lv[5] == 23928800
    with torch.no_grad():
        ref.params.mul_(1e4)  # O(1) table values so that fp16 rounding is exercised
        enc.params.copy_(ref.params)
    g = torch.Generator().manual_seed(5)
    n = 4099
    x = torch.rand(n, 3, generator=g)
    x[:5] = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.5, 0.5, 0.5], [1.0, 0.0, 0.25], [-0.2, 1.3, 0.5]])
    gy = torch.randn(n, 2 * c["n_levels"], generator=g).half().float()  # the encoder output is fp16, so is dy
    xr = x.clone().requires_grad_(True)
    y_ref = ref(xr)
    (y_ref * gy).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    y = enc(xg)
    assert y.dtype == torch.float16
    (y.float() * gy.to(DEV)).sum().backward()
    inside = ((x >= 0) & (x <= 1)).all(-1)
    assert torch.equal(y.float().cpu()[inside], y_ref.detach()[inside])  # same op order -> same fp16 bits
    torch.testing.assert_close(enc.params.grad.cpu(), ref.params.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(xg.grad.cpu()[inside], xr.grad[inside], rtol=1e-3, atol=1e-3 * float(xr.grad.abs().max()))
    # linearity in the table (size-independent property)
    with torch.no_grad():
        enc.params.mul_(2.0)
    y2 = enc(x.to(DEV)).float()
    torch.testing.assert_close(y2, 2 * y.detach().float(), rtol=2e-3, atol=1e-4)


@pytest.mark.parametrize("compat", [False, True])
def test_hashgrid4d_fwd_bwd(cb, compat):
    ref = tr.HashEncoder4D(max_params=2 ** 14, levels=8, base_res=16.0, max_res=256.0, taichi_compat=compat, seed=3)
    enc = cb.hash_encoder.HashEncoder4D(max_params=2 ** 14, levels=8, base_res=16.0, max_res=256.0,
                                        taichi_compat=compat, seed=3).to(DEV)
    with torch.no_grad():
        ref.hash_table.mul_(1e4)
        enc.hash_table.copy_(ref.hash_table)
    g = torch.Generator().manual_seed(9)
    n = 3001
    x = torch.rand(n, 4, generator=g)
    x[:4, 3] = torch.tensor([0.0, 1.0, 1.0 / 3.0, 2.0 / 3.0])
    gy = torch.randn(n, 16, generator=g).half().float()
    y_ref = ref(x)
    (y_ref * gy).sum().backward()
    y = enc(x.to(DEV))
    (y.float() * gy.to(DEV)).sum().backward()
    assert torch.equal(y.float().cpu(), y_ref.detach())
    torch.testing.assert_close(enc.hash_table.grad.cpu(), ref.hash_table.grad, rtol=1e-3, atol=1e-5)


def test_hash_encoder_taichi_surface(cb):
    ref = tr.HashEncoder(max_params=2 ** 12, levels=8, base_res=16.0, max_res=256.0, seed=3)
    enc = cb.hash_encoder.HashEncoder(max_params=2 ** 12, levels=8, base_res=16.0, max_res=256.0, seed=3).to(DEV)
    assert enc.begin_fast_hash_level == ref.begin_fast_hash_level and enc.out_dim == 16
    assert torch.equal(enc.offsets.cpu(), ref.offsets) and torch.equal(enc.hash_map_sizes.cpu(), ref.hash_map_sizes)
    with torch.no_grad():
        enc.hash_table.copy_(ref.hash_table)
    x = torch.rand(1000, 3, generator=torch.Generator().manual_seed(1))
    assert torch.equal(enc(x.to(DEV)).float().cpu(), ref(x).detach())


# ---------------------------------------------------------------------------------------------- encodings
def test_frequency_sh_time(cb):
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(2000, 4, generator=g) * 2 - 1) * 1.5
    gy = torch.randn(2000, 32, generator=g).half().float()
    xr = x.clone().requires_grad_(True)
    y_ref = tc.frequency_encode(xr, 4)
    (y_ref * gy).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    y = cb.tcnn.Encoding(4, {"otype": "Frequency", "n_frequencies": 4})(xg)
    (y.float() * gy.to(DEV)).sum().backward()
    torch.testing.assert_close(y.float().cpu(), y_ref.detach(), rtol=0, atol=1e-3)  # one fp16 ulp near 1
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-3)
    d = torch.rand(2000, 3, generator=g)
    torch.testing.assert_close(cb.ops.sh2_encode(d.to(DEV)).float().cpu(), tc.sh_encode_deg2(d), rtol=0, atol=5e-4)
    from oracle import cednerf_ref as cr

    t, mv = torch.rand(500, 1, generator=g), torch.rand(500, 1, generator=g) * 0.01
    torch.testing.assert_close(cb.encoder.SinusoidalEncoder(1, 0, 4, True)(t.to(DEV)).cpu(), cr.time_embed(t),
                               rtol=0, atol=1e-6)
    torch.testing.assert_close(cb.encoder.SinusoidalEncoderWithExp(1, 0, 4, True)(t.to(DEV), mv.to(DEV)).cpu(),
                               cr.time_embed_attenuated(t, mv), rtol=0, atol=1e-6)


# ---------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n_in,n_out,n_hidden,n", [(32, 6, 3, 1000), (41, 16, 1, 4096), (19, 3, 2, 777), (32, 32, 1, 130),
                                                    (32, 1, 1, 128), (64, 64, 4, 300)])
def test_mlp_fwd_bwd(cb, n_in, n_out, n_hidden, n):
    cfg = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64,
           "n_hidden_layers": n_hidden}
    ref = tc.Network(n_in, n_out, cfg, seed=5)
    net = cb.tcnn.Network(n_in, n_out, cfg, seed=5).to(DEV)
    assert net.params.numel() == ref.params.numel()
    with torch.no_grad():
        net.params.copy_(ref.params)
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(n, n_in, generator=g)).half().float()
    gy = torch.randn(n, n_out, generator=g)
    xr = x.clone().requires_grad_(True)
    y_ref = ref(xr)
    (y_ref * gy).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    y = net(xg)
    assert y.dtype == torch.float16 and y.shape == (n, n_out)
    (y.float() * gy.to(DEV)).sum().backward()
    scale = float(y_ref.detach().abs().max())
    torch.testing.assert_close(y.float().cpu(), y_ref.detach(), rtol=2e-3, atol=2e-3 * scale)  # fp16 MLP path <= 2e-3
    gp, gp_ref = net.params.grad.cpu(), ref.params.grad
    assert float((gp - gp_ref).norm() / gp_ref.norm()) < 1e-3                               # gradient rel. error <= 1e-3
    torch.testing.assert_close(gp, gp_ref, rtol=1e-2, atol=2e-3 * float(gp_ref.abs().max()))
    gx, gx_ref = xg.grad.cpu(), xr.grad
    assert float((gx - gx_ref).norm() / gx_ref.norm()) < 1e-3
    # inference path (no saved activations) gives the same output
    with torch.no_grad():
        assert torch.equal(net(x.to(DEV)), y.detach())


def test_mlp_with_input_encoding(cb):
    cfg = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64,
           "n_hidden_layers": 3}
    freq = {"otype": "Frequency", "n_frequencies": 4}
    ref = tc.NetworkWithInputEncoding(4, 6, freq, cfg, seed=2)
    net = cb.tcnn.NetworkWithInputEncoding(4, 6, freq, cfg, seed=2).to(DEV)
    with torch.no_grad():
        net.params.copy_(ref.params)
    x = torch.rand(5000, 4, generator=torch.Generator().manual_seed(4)) * 2 - 1
    y_ref = ref(x).detach()
    y = net(x.to(DEV)).float().cpu()
    torch.testing.assert_close(y, y_ref, rtol=2e-3, atol=2e-3 * float(y_ref.abs().max()))


@pytest.mark.parametrize("reduce", ["sum", "mean"])
@pytest.mark.parametrize("weighted", [False, True])
def test_reduce_along_rays_against_scatter_reduce(cb, reduce, weighted):
    """cednerf/render.py:8-39 is torch scatter_reduce_ with include_self=True; here a segmented kernel (+ the count
    division for 'mean'), forward and backward to both values and weights."""
    g = torch.Generator().manual_seed(12)
    n_rays, c = 700, 5
    counts = torch.randint(0, 40, (n_rays,), generator=g)
    counts[::9] = 0
    ridx = torch.repeat_interleave(torch.arange(n_rays), counts)
    v = torch.randn(ridx.numel(), c, generator=g)
    w = torch.rand(ridx.numel(), 1, generator=g) if weighted else None
    v_ref = v.clone().double().requires_grad_(True)
    w_ref = None if w is None else w.clone().double().requires_grad_(True)
    src = v_ref if w_ref is None else w_ref * v_ref
    want = torch.zeros(n_rays, c, dtype=torch.float64).scatter_reduce(0, ridx[:, None].expand(-1, c), src, reduce=reduce)
    go = torch.randn(n_rays, c, generator=g)
    want.backward(go.double())
    v_g = v.to(DEV).requires_grad_(True)
    w_g = None if w is None else w.to(DEV).requires_grad_(True)
    got = cb.render.reduce_along_rays(ridx.to(DEV), v_g, n_rays, w_g, reduce)
    got.backward(go.to(DEV))
    torch.testing.assert_close(got.detach().cpu().double(), want.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(v_g.grad.cpu().double(), v_ref.grad, rtol=1e-5, atol=1e-6)
    if weighted:
        torch.testing.assert_close(w_g.grad.cpu().double(), w_ref.grad, rtol=1e-5, atol=1e-6)
