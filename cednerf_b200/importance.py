"""Importance-sampled training batches: the training branch of the reference's SubjectLoader.fetch_data
(datasets/dnerf_3d_video_IS.py:401-497) on the device, with no host read.

    sampler = ImportanceSampler(images_u8, camtoworlds, K, timestamps, sampling_weights, weights_subsampled=2)
    sampler.update_num_rays(n)            # dnerf_3d_video_IS.py:398
    data = sampler.fetch_data()           # {"rgb", "rays", "timestamps", "idx"} as the reference returns them

`sampling_weights` are the ISG / IST maps the reference loads or computes once per dataset (:225-262) - one weight per
cell of the `weights_subsampled`-times down-sampled frames of every image; building them is dataset preparation and stays
with the caller.  What runs every step is here: thin the weights to a uniform random subset (torch.randint), draw
`num_rays / s^2` cells WITHOUT replacement in proportion to their weights, expand every cell to its s x s pixels, gather
their colours and generate their rays.

torch.multinomial(w, k) without replacement is `topk(w / Exp(1) noise, k)` (ATen's multinomial: exponential_(1), div,
topk); `weighted_sample` is that selection as a radix select + ordered compaction (csrc/importance.cu).  It returns the
winners in ascending position order (torch returns them by descending key; a batch is a set).  Passing `noise` makes the
draw a pure function of its inputs - the parity tests hand the same noise to the oracle."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops
from ._lib import call, ptr, stream
from .utils import Rays


def weighted_sample(weights: torch.Tensor, k: int, subset: Optional[torch.Tensor] = None,
                    noise: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None,
                    check: bool = False) -> torch.Tensor:
    """k distinct positions of `weights` (or of weights[subset], returned as entries of `subset`) drawn in proportion to
    the weights, without replacement.  noise: Exp(1) draws, one per candidate (default: torch's exponential_ on the
    device).  check=True reads the error flag back (fewer than k positive weights -> RuntimeError, as torch raises)."""
    _lib.check_device()
    w = ops._f32c(weights).view(-1)
    dev = w.device
    sub = None if subset is None else subset.to(dev, torch.int64).contiguous().view(-1)
    n = w.numel() if sub is None else sub.numel()
    k = int(k)
    if k > n:
        raise RuntimeError("cannot sample n_sample > prob_dist.size(-1) samples without replacement")
    if noise is None:
        noise = torch.empty(n, device=dev).exponential_(1.0, generator=generator)
    noise = ops._f32c(noise).view(-1)
    if noise.numel() != n:
        raise ValueError("one noise draw per candidate")
    keys = torch.empty(n, dtype=torch.int32, device=dev)
    out = torch.empty(k, dtype=torch.int64, device=dev)
    if k == 0:
        return out
    call("cednerf_importance_keys", ptr(w), ptr(sub), ptr(noise), n, ptr(keys), stream())
    ws = torch.empty(int(_lib.load().cednerf_topk_workspace_bytes(n)) // 8 + 1, dtype=torch.int64, device=dev)
    call("cednerf_topk_select", ptr(keys), n, k, ptr(sub), ptr(out), ptr(ws), stream())
    if check and int(ws.view(torch.int32)[4]) != 0:
        raise RuntimeError("invalid multinomial distribution (with replacement=False, not enough non-negative category to sample)")
    return out


class ImportanceSampler:
    """The per-step half of datasets/dnerf_3d_video_IS.py::SubjectLoader for training (constructor arguments are the
    attributes fetch_data reads).  images: uint8 [n_images, H, W, 3]; camtoworlds: [n_images, 3|4, 4]; K: 3 x 3;
    timestamps: [n_images, 1]; sampling_weights: [n_images * (H // s) * (W // s)]."""

    def __init__(self, images, camtoworlds, K, timestamps, sampling_weights, weights_subsampled: int = 1,
                 sampling_batch_size: int = 2_000_000, num_rays: int = 4096, opengl_camera: bool = False,
                 generator: Optional[torch.Generator] = None):
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError("images: uint8 [n_images, H, W, 3]")
        self.images = images.contiguous()
        self.camtoworlds = ops._f32c(camtoworlds)
        Kh = torch.as_tensor(K, dtype=torch.float32).cpu()
        self.K = Kh
        self._intr = (float(Kh[0, 0]), float(Kh[1, 1]), float(Kh[0, 2]), float(Kh[1, 2]))
        self.timestamps = ops._f32c(timestamps).view(-1, 1)
        self.sampling_weights = ops._f32c(sampling_weights).view(-1)
        self.weights_subsampled = int(weights_subsampled)
        self.sampling_batch_size = int(sampling_batch_size)
        self.num_rays = int(num_rays)
        self.OPENGL_CAMERA = bool(opengl_camera)
        self.generator = generator
        self.height, self.width = int(images.shape[1]), int(images.shape[2])
        s = self.weights_subsampled
        if self.sampling_weights.numel() != images.shape[0] * (self.height // s) * (self.width // s):
            raise ValueError("one weight per cell of the subsampled frames")
        self.training = True

    def update_num_rays(self, num_rays):
        self.num_rays = int(num_rays)

    def draw_cells(self, subset=None, noise=None):
        """The multinomial draw of :405-418 -> int64 [num_rays // s^2] cell indices."""
        batch_size = self.num_rays // (self.weights_subsampled ** 2)
        n_w = self.sampling_weights.numel()
        if n_w > self.sampling_batch_size and subset is None:
            subset = torch.randint(0, n_w, (self.sampling_batch_size,), dtype=torch.int64,
                                   device=self.sampling_weights.device, generator=self.generator)
        return weighted_sample(self.sampling_weights, batch_size, subset=subset, noise=noise, generator=self.generator)

    def fetch_data(self, subset=None, noise=None):
        cells = self.draw_cells(subset, noise)
        k, s = cells.numel(), self.weights_subsampled
        n, dev = k * s * s, cells.device
        o, d, rgb = (torch.empty(n, 3, device=dev) for _ in range(3))
        ts = torch.empty(n, 1, device=dev)
        image_id = torch.empty(n, dtype=torch.int64, device=dev)
        fx, fy, cx, cy = self._intr
        call("cednerf_importance_batch", ptr(cells), k, s, self.width, self.height, ptr(self.images), ptr(self.camtoworlds),
             int(self.camtoworlds.shape[1]), fx, fy, cx, cy, int(self.OPENGL_CAMERA), ptr(self.timestamps), ptr(o), ptr(d),
             ptr(rgb), ptr(ts), ptr(image_id), None, stream())
        return {"rgb": rgb, "rays": Rays(origins=o, viewdirs=d), "timestamps": ts, "idx": image_id}
