"""Synthetic, dataset-free workloads shaped like the reference's configurations.

Nothing here is on the hot path: it builds the inputs the benchmark, the smoke test and the full-size
property tests feed to it (cameras -> rays, analytic occupancy, random-init weights with a density boost).
All draws come from seeded CPU generators, so the GPU path and the CPU oracle can be given identical inputs."""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field
from typing import Dict, Tuple

import torch


@dataclass
class SceneConfig:
    name: str
    width: int
    height: int
    focal: float
    n_cams: int
    n_frames: int
    roi_aabb: Tuple[float, ...]
    occ_levels: int
    occ_res: int
    near_plane: float
    render_step_size: float
    cone_angle: float
    alpha_thre: float
    dst_resolution: int
    log2_hashmap_size: int
    moving_step: float
    per_ray_time: bool
    flags: Dict[str, bool] = dc_field(default_factory=dict)


# train_real.py:151-182 / train_prop_real.py:160-191 (factor 2), run_dynerf.sh flags -te -ta -df -f -wr -ae
DYNERF = SceneConfig("dynerf_flame_salmon_1_shaped", 1352, 1014, 731.0, 18, 300, (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0), 4, 128,
                     0.2, 1e-3, 0.004, 1e-2, 8192, 21, 1.0 / 8192, True,
                     dict(use_time_embedding=True, use_time_attenuation=True, use_div_offsets=True, use_feat_predict=True))
# train_real.py:119-149, run_hyper.sh flags -te -ta -f -ae -df -d: one camera / one timestamp per batch
HYPERNERF = SceneConfig("hypernerf_vrig_3dprinter_shaped", 536, 960, 700.0, 1, 300, (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0), 2,
                        128, 0.2, 1e-3, 0.004, 1e-2, 4096, 21, 1.0 / 4096, False,
                        dict(use_time_embedding=True, use_time_attenuation=True, use_div_offsets=True,
                             use_feat_predict=True))
# train_real.py:86-117: 800x800, camera_angle_x 0.6911 -> focal 1111.1, radius-4 orbit
DNERF = SceneConfig("dnerf_synthetic_shaped", 800, 800, 1111.1, 1, 100, (-1.5, -1.5, -1.5, 1.5, 1.5, 1.5), 1, 128, 0.0, 5e-3,
                    0.0, 0.0, 1024, 21, 1e-4, True, dict(use_time_embedding=True, use_time_attenuation=True))
# small stand-in used by CPU-side tests and the smoke test
TINY = SceneConfig("tiny", 64, 48, 40.0, 3, 8, (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0), 2, 16, 0.2, 2e-2, 0.004, 1e-2, 256, 12,
                   1.0 / 256, True, dict(use_time_embedding=True, use_time_attenuation=True, use_div_offsets=True,
                                         use_feat_predict=True))


def blob_occupancy(cfg: SceneConfig, seed: int = 42, n_blobs: int = 3, radius_frac: float = 0.5) -> torch.Tensor:
    """bool [L,R,R,R]: a cell is occupied iff its box intersects one of `n_blobs` spheres (radius = radius_frac x
    half ROI extent) at seeded centres; level l covers centre +- half*2^l (nerfacc's nested grids)."""
    g = torch.Generator().manual_seed(seed)
    roi = torch.tensor(cfg.roi_aabb)
    centre, half = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
    blobs = centre + (torch.rand(n_blobs, 3, generator=g) - 0.5) * half
    radius = radius_frac * float(half.min())
    r = cfg.occ_res
    out = torch.zeros(cfg.occ_levels, r, r, r, dtype=torch.bool)
    idx = torch.arange(r, dtype=torch.float32)
    for l in range(cfg.occ_levels):
        lo = centre - half * 2 ** l
        cell = (half * 2 ** l) * 2 / r
        for b in blobs:
            d2 = torch.zeros(r, r, r)
            for ax in range(3):
                c_lo = lo[ax] + idx * cell[ax]
                gap = torch.clamp(torch.maximum(c_lo - b[ax], b[ax] - (c_lo + cell[ax])), min=0.0) ** 2
                shape = [1, 1, 1]
                shape[ax] = r
                d2 = d2 + gap.view(shape)
            out[l] |= d2 <= radius * radius
    return out


def camera_centres(cfg: SceneConfig) -> torch.Tensor:
    """Forward-facing rig (OpenCV convention, looking down +z) on a 0.4-wide arc in front of the ROI."""
    roi = torch.tensor(cfg.roi_aabb)
    half = float((roi[3:] - roi[:3]).min()) / 2
    xs = torch.linspace(-0.2, 0.2, cfg.n_cams) if cfg.n_cams > 1 else torch.zeros(1)
    return torch.stack([xs * half, 0.05 * half * torch.sin(xs * 7.0), torch.full_like(xs, -1.6 * half)], -1)


def draw_batch(cfg: SceneConfig, n_rays: int, gen: torch.Generator, pin: bool = False):
    """One training batch on the HOST: rays drawn uniformly over (camera, frame, pixel) (dnerf_3d_video_IS.py
    batching), U(0,1) pixels, random background.  -> dict of CPU tensors (pinned if asked)."""
    cams = camera_centres(cfg)
    cam = torch.randint(0, cfg.n_cams, (n_rays,), generator=gen)
    u = torch.randint(0, cfg.width, (n_rays,), generator=gen).float()
    v = torch.randint(0, cfg.height, (n_rays,), generator=gen).float()
    d = torch.stack([(u - cfg.width / 2 + 0.5) / cfg.focal, (v - cfg.height / 2 + 0.5) / cfg.focal, torch.ones(n_rays)], -1)
    d = d / torch.linalg.norm(d, dim=-1, keepdim=True)
    if cfg.per_ray_time:
        t = torch.randint(0, cfg.n_frames, (n_rays, 1), generator=gen).float() / max(cfg.n_frames - 1, 1)
    else:
        t = (torch.randint(0, cfg.n_frames, (1, 1), generator=gen).float() / max(cfg.n_frames - 1, 1)).expand(n_rays, 1).contiguous()
    batch = {"origins": cams[cam].contiguous(), "viewdirs": d.contiguous(), "timestamps": t,
             "pixels": torch.rand(n_rays, 3, generator=gen), "color_bkgd": torch.rand(3, generator=gen),
             "jitter": torch.rand(n_rays, generator=gen)}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def frame_rays(cfg: SceneConfig, cam_index: int = 0, rows=None):
    """All rays of one frame (or of the row block `rows` = (lo, hi)) on the host, row-major."""
    lo, hi = (0, cfg.height) if rows is None else rows
    v, u = torch.meshgrid(torch.arange(lo, hi).float(), torch.arange(cfg.width).float(), indexing="ij")
    d = torch.stack([(u - cfg.width / 2 + 0.5) / cfg.focal, (v - cfg.height / 2 + 0.5) / cfg.focal, torch.ones_like(u)], -1)
    d = (d / torch.linalg.norm(d, dim=-1, keepdim=True)).reshape(-1, 3)
    o = camera_centres(cfg)[cam_index].expand_as(d).contiguous()
    return o, d.contiguous()


def boost_density(field) -> None:
    """Random-init sigma is ~exp(-1) everywhere, which the alpha threshold would filter out entirely.  Widen the
    logits (the recipe of tests/golden/make_golden.py, table x20000 -> U(-2,2)) so that a realistic share of the
    marched samples survives: on the DyNeRF-shaped scene ~35 marched and ~3.9 surviving samples per ray, i.e.
    ~2^20 samples for 2^18 rays - the reference's target_sample_batch_size (train_real.py:157)."""
    with torch.no_grad():
        field.hash_encoder.params.mul_(20000.0)
        field.mlp_base.params.mul_(3.0)


def field_kwargs(cfg: SceneConfig) -> dict:
    return dict(dst_resolution=cfg.dst_resolution, log2_hashmap_size=cfg.log2_hashmap_size, moving_step=cfg.moving_step,
                **cfg.flags)


def render_kwargs(cfg: SceneConfig) -> dict:
    return dict(near_plane=cfg.near_plane, render_step_size=cfg.render_step_size, cone_angle=cfg.cone_angle,
                alpha_thre=cfg.alpha_thre)


def initial_state(cfg: SceneConfig, seed: int = 1337) -> Dict[str, torch.Tensor]:
    """The random initialisation of DNGPradianceField(**field_kwargs(cfg), seed=seed) as a plain state dict, written with
    torch only (hash table U(-1e-4, 1e-4), Xavier-uniform [out, in] layers on the padded shapes, seeded CPU generators -
    the arithmetic of cednerf_b200/tcnn.py, checked against it in tests/test_cpu_abi.py).  Both benchmark arms load it, so
    the CPU restatement and the CUDA path start from bit-identical weights without importing one another."""
    kw = field_kwargs(cfg)
    n_levels, base = 16, 16
    per_level = math.exp(math.log(kw["dst_resolution"] / base) / (n_levels - 1))
    total = 0
    for l in range(n_levels):
        res = int(math.ceil(base * math.exp(l * math.log(per_level)) - 1.0)) + 1
        total += min(2 ** kw["log2_hashmap_size"], (res ** 3 + 7) // 8 * 8)

    def pad16(n):
        return (n + 15) // 16 * 16

    def mlp(n_in, n_out, n_hidden, sd):
        g = torch.Generator().manual_seed(sd)
        dims = [pad16(n_in)] + [64] * n_hidden + [pad16(n_out)]
        return torch.cat([((torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (o + i))).reshape(-1)
                          for i, o in zip(dims[:-1], dims[1:])])

    te, ta = kw.get("use_time_embedding", False), kw.get("use_time_attenuation", False)
    before = kw.get("time_inject_before_sigma", True)
    g = torch.Generator().manual_seed(seed + 2)
    roi = torch.tensor(cfg.roi_aabb)
    centre, half = (roi[:3] + roi[3:]) / 2, (roi[3:] - roi[:3]) / 2
    scale = 2 ** (cfg.occ_levels - 1)
    sd = {"aabb": torch.cat([centre - half * scale, centre + half * scale]),
          "xyz_wrap.params": mlp(32, 6 if kw.get("use_div_offsets") else 3, 3, seed + 1),
          "direction_encoding.params": torch.zeros(0),
          "hash_encoder.params": (torch.rand(total * 2, generator=g) * 2 - 1) * 1e-4,
          "mlp_base.params": mlp(2 * n_levels + (9 if te and before else 0), 16, 1, seed + 3),
          "mlp_head.params": mlp(4 + 15 + (9 if te and not before else 0), 3, 2, seed + 4)}
    if te:
        sd["time_encoder.scales"] = torch.tensor([1, 2, 4, 8])
        sd["time_encoder_feat.scales"] = torch.tensor([1, 2, 4, 8])
        sd["time_encoder_feat.scales_move"] = torch.tensor([0, 2, 8, 24])
    if kw.get("use_feat_predict"):
        sd["mlp_feat_prediction.params"] = mlp(32, 2 * n_levels, 1, seed + 5)
    if kw.get("use_weight_predict"):
        sd["mlp_weight_prediction.params"] = mlp(32, 1, 1, seed + 6)
    return sd


def build_scene(cfg: SceneConfig, device, impl, seed: int = 42):
    """(estimator, field) for `impl` in {cednerf_b200, oracle.cednerf_ref-like namespace}: same occupancy, same
    initial weights (`initial_state`), same density boost."""
    est = impl.OccGridEstimator(list(cfg.roi_aabb), resolution=cfg.occ_res, levels=cfg.occ_levels)
    binaries = blob_occupancy(cfg, seed)
    fld = impl.DNGPradianceField(est.aabbs[-1], **field_kwargs(cfg))
    fld.load_state_dict(initial_state(cfg))
    boost_density(fld)
    est, fld = est.to(device), fld.to(device)
    est.binaries = binaries.to(device)
    est.occs = binaries.flatten().float().to(device) * 0.5
    return est, fld


# ---- render poses (BASELINE.json configs[0] and [3]) -------------------------------------------------------------------
def _normalize(v):
    return v / torch.linalg.norm(v)


def _viewmatrix(z, up, pos):
    """datasets/utils.py:39-46 (camera-to-world from a viewing axis, an up hint and a position)."""
    x = _normalize(torch.linalg.cross(up, z))
    y = torch.linalg.cross(z, x)
    return torch.stack([x, y, z, pos], 1)


def spiral_poses(cfg: SceneConfig, n_frames: int = 300, n_rots: int = 2, zrate: float = 0.5, dt: float = 0.75,
                 percentile: float = 70.0) -> torch.Tensor:
    """[n_frames, 3, 4] camera-to-world matrices of the novel-view video (datasets/utils.py:67-112, generate_spiral_path)
    around the synthetic forward-facing rig of `camera_centres`: spiral radii from the 70th percentile of the camera
    offsets, all poses looking at one focus point in front of the rig.  OpenCV convention (camera looks down +z)."""
    cams = camera_centres(cfg).double()
    centre = cams.mean(0)
    half = float((torch.tensor(cfg.roi_aabb[3:]) - torch.tensor(cfg.roi_aabb[:3])).min()) / 2
    close_depth, inf_depth = 0.6 * half, 2.6 * half * 5.0
    focal = 1.0 / ((1.0 - dt) / close_depth + dt / inf_depth)
    radii = torch.quantile((cams - centre).abs(), percentile / 100.0, dim=0).clamp_min(0.02 * half)
    up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)  # image y points down in the OpenCV convention
    lookat = centre + torch.tensor([0.0, 0.0, focal], dtype=torch.float64)
    poses = []
    for k in range(n_frames):
        th = 2.0 * math.pi * n_rots * k / n_frames
        pos = centre + radii * torch.tensor([math.cos(th), -math.sin(th), -math.sin(th * zrate)], dtype=torch.float64)
        poses.append(_viewmatrix(_normalize(lookat - pos), -up, pos))
    return torch.stack(poses).float()


def orbit_pose(radius: float, theta: float, phi: float = 0.5) -> torch.Tensor:
    """[3, 4] camera on a sphere of `radius` looking at the origin, OpenGL convention (camera looks down -z), as the
    D-NeRF synthetic cameras (dnerf_synthetic.py:202-221)."""
    pos = radius * torch.tensor([math.cos(phi) * math.cos(theta), math.cos(phi) * math.sin(theta), math.sin(phi)],
                                dtype=torch.float64)
    z = _normalize(pos)  # camera z axis points away from the scene
    return _viewmatrix(z, torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64), pos).float()


def pose_rays(cfg: SceneConfig, c2w: torch.Tensor, opengl: bool = False, device=None):
    """All rays of one frame seen from `c2w` [3,4] -> (origins [H*W,3], unit viewdirs [H*W,3]), row-major, built on
    `device` (dnerf_3d_video_IS.py:340-358 OpenCV / dnerf_synthetic.py:202-221 OpenGL pixel -> ray)."""
    c2w = c2w.to(device) if device is not None else c2w
    v, u = torch.meshgrid(torch.arange(cfg.height, device=c2w.device, dtype=torch.float32),
                          torch.arange(cfg.width, device=c2w.device, dtype=torch.float32), indexing="ij")
    x, y = (u - cfg.width / 2 + 0.5) / cfg.focal, (v - cfg.height / 2 + 0.5) / cfg.focal
    cam = torch.stack([x, -y, -torch.ones_like(x)], -1) if opengl else torch.stack([x, y, torch.ones_like(x)], -1)
    d = (cam.reshape(-1, 3) @ c2w[:, :3].T)
    d = d / torch.linalg.norm(d, dim=-1, keepdim=True)
    return c2w[:, 3].expand_as(d).contiguous(), d.contiguous()
