"""GPU parity of OccGridEstimator.mark_invisible_cells (nerfacc call of train_real.py:205-211) against the oracle's
restatement: occs are exactly 0 / -1, compared bit for bit (same fp32 operations in the same order); `_update` then
leaves the invisible cells alone, as nerfacc does (occs >= 0 masks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cameras(seed, n, rows):
    g = torch.Generator().manual_seed(seed)
    c2w = torch.zeros(n, 4, 4)
    for i in range(n):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        c2w[i, :3, :3] = q
        c2w[i, :3, 3] = (torch.rand(3, generator=g) - 0.5) * 6
        c2w[i, 3, 3] = 1
    return c2w[:, :rows].contiguous()


@pytest.mark.parametrize("n_cams,rows,per_cam_K,res,levels", [(3, 4, False, 16, 2), (70, 3, True, 32, 1), (1, 4, False, 8, 4)])
def test_mark_invisible_cells_matches_oracle(n_cams, rows, per_cam_K, res, levels):
    import cednerf_b200 as cb
    from oracle import nerfacc_ref as nf

    c2w = _cameras(5 + n_cams, n_cams, rows)
    K = torch.tensor([[40.0, 0.0, 32.0], [0.0, 42.0, 24.0], [0.0, 0.0, 1.0]])[None]
    if per_cam_K:
        K = K.repeat(n_cams, 1, 1)
        K[:, 0, 0] += torch.arange(n_cams) * 0.5
    ref = nf.OccGridEstimator([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], resolution=res, levels=levels)
    ref.mark_invisible_cells(K, c2w, 64, 48, 0.3)
    est = cb.OccGridEstimator([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], resolution=res, levels=levels).to(DEV)
    launches = cb._lib.launch_count()
    est.mark_invisible_cells(K.to(DEV), c2w.to(DEV), 64, 48, 0.3)
    assert cb._lib.launch_count() == launches + 1
    assert torch.equal(est.occs.cpu(), ref.occs)
    frac = float((ref.occs < 0).float().mean())
    assert 0.0 < frac < 1.0 or n_cams == 1, frac   # the case is not degenerate

    # an occupancy update afterwards never revives an invisible cell
    est.train()
    est._update(0, lambda x: torch.full((x.shape[0], 1), 5.0, device=x.device))
    occs = est.occs.cpu()
    assert torch.equal(occs < 0, ref.occs < 0)
    assert not bool(est.binaries.flatten().cpu()[ref.occs < 0].any())
    assert bool(est.binaries.flatten().cpu()[ref.occs >= 0].all())
