"""GPU parity of the importance-sampled batch (csrc/importance.cu, cednerf_b200/importance.py) against the restated
fetch_data of datasets/dnerf_3d_video_IS.py:401-497 (oracle/dataset_ref.py, itself pinned to torch.multinomial on the
CPU): the SET of drawn cells is bit-exact given the same subset and Exp(1) draws; colours exact; rays within 2e-7."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def scene(seed, n_images, H, W, s):
    g = torch.Generator().manual_seed(seed)
    images = torch.randint(0, 256, (n_images, H, W, 3), generator=g, dtype=torch.uint8)
    ang = torch.rand(n_images, generator=g) * 0.6 - 0.3
    c2w = torch.zeros(n_images, 3, 4)
    c2w[:, 0, 0], c2w[:, 0, 2], c2w[:, 2, 0], c2w[:, 2, 2] = torch.cos(ang), torch.sin(ang), -torch.sin(ang), torch.cos(ang)
    c2w[:, 1, 1] = 1.0
    c2w[:, :, 3] = torch.rand(n_images, 3, generator=g) - 0.5
    K = torch.tensor([[0.9 * W, 0, W / 2], [0, 0.9 * W, H / 2], [0, 0, 1.0]])
    ts = torch.linspace(0, 1, n_images)[:, None].contiguous()
    weights = torch.rand(n_images * (H // s) * (W // s), generator=g) ** 6   # heavy-tailed, like ISG maps
    weights[torch.rand(weights.numel(), generator=g) < 0.3] = 0.0
    return images, c2w, K, ts, weights, g


@pytest.mark.parametrize("n_images,H,W,s,num_rays,subset_size", [(6, 48, 64, 2, 4096, None), (5, 40, 60, 1, 1000, 4000),
                                                                 (40, 96, 128, 4, 16384, 20000), (3, 16, 16, 2, 8, None)])
def test_importance_batch_matches_the_restated_fetch_data(n_images, H, W, s, num_rays, subset_size):
    import cednerf_b200 as cb
    from oracle import dataset_ref as dr

    images, c2w, K, ts, weights, g = scene(n_images + H, n_images, H, W, s)
    n_w = weights.numel()
    subset = None if subset_size is None else torch.randint(0, n_w, (subset_size,), generator=g)
    noise = torch.empty(n_w if subset is None else subset_size).exponential_(1, generator=g)
    want = dr.fetch_data_train(images, c2w, K, ts, weights, s, num_rays, W, H, False, subset, noise)
    sampler = cb.importance.ImportanceSampler(images.to(DEV), c2w.to(DEV), K, ts.to(DEV), weights.to(DEV), s,
                                              sampling_batch_size=10 ** 9 if subset is None else subset_size, num_rays=num_rays)
    launches = cb._lib.launch_count()
    got = sampler.fetch_data(subset=None if subset is None else subset.to(DEV), noise=noise.to(DEV))
    assert cb._lib.launch_count() - launches == 12   # keys, 10 of the select, batch assembly
    k = num_rays // (s * s)
    # the batch is the reference's up to the order of the drawn cells: ours ascending, torch's by descending key
    cells_got = sampler.draw_cells(subset=None if subset is None else subset.to(DEV), noise=noise.to(DEV)).cpu()
    if subset is None:
        assert cells_got.unique().numel() == k                      # without replacement
    order_w = torch.argsort(want["cells"], stable=True)
    assert torch.equal(torch.sort(cells_got)[0], want["cells"][order_w])
    if subset is None:
        assert torch.equal(cells_got, want["cells"][order_w])        # ascending position order
    # match rays by pixel: (image, y, x) is a unique key of a ray when cells are distinct
    if subset is None:
        perm = torch.cat([order_w + j * k for j in range(s * s)])
        for key in ("rgb", "origins", "viewdirs", "timestamps", "idx"):
            a = {"origins": got["rays"].origins, "viewdirs": got["rays"].viewdirs}.get(key, got.get(key)).cpu()
            b = want[key][perm]
            if key == "viewdirs":
                torch.testing.assert_close(a, b, rtol=0, atol=2e-7)
            else:
                assert torch.equal(a, b), key


def test_weighted_sample_properties_at_full_size():
    """2 M candidates (the reference's sampling_batch_size), 16 384 draws: distinct, positive-weight only, equal to
    torch.topk's set; zero-weight shortage raises with check=True; k = n returns everything."""
    import cednerf_b200 as cb

    g = torch.Generator().manual_seed(3)
    n, k = 2_000_000, 16384
    w = (torch.rand(n, generator=g) ** 8).to(DEV)
    w[::3] = 0
    noise = torch.empty(n, device=DEV).exponential_(1)
    got = cb.importance.weighted_sample(w, k, noise=noise, check=True)
    want = torch.sort(torch.topk(w / noise, k).indices)[0]
    assert torch.equal(got, want) and bool((w[got] > 0).all())
    few = torch.zeros(1000, device=DEV)
    few[:10] = 1.0
    with pytest.raises(RuntimeError):
        cb.importance.weighted_sample(few, 11, check=True)
    assert torch.equal(cb.importance.weighted_sample(few, 10, check=True), torch.arange(10, device=DEV))
    allw = torch.rand(777, generator=g).to(DEV) + 0.1
    assert torch.equal(cb.importance.weighted_sample(allw, 777), torch.arange(777, device=DEV))
    # draws follow the weights: two classes with 3 : 1 weights, 25 % of the population drawn
    cls = torch.ones(400_000, device=DEV)
    cls[::2] = 3.0
    picked = cb.importance.weighted_sample(cls, 100_000)
    frac_heavy = float((cls[picked] == 3.0).float().mean())
    assert 0.69 < frac_heavy < 0.73     # exact value for sampling without replacement at this depth: ~0.711
