"""cednerf/encoder.py surface: the two 1-D sinusoidal time encoders (cednerf/encoder.py:6-44, :46-90)."""
from __future__ import annotations

import torch

from . import ops


class SinusoidalEncoder(torch.nn.Module):
    def __init__(self, x_dim=1, min_deg=0, max_deg=4, use_identity=True):
        super().__init__()
        if (x_dim, min_deg, max_deg, use_identity) != (1, 0, 4, True):
            raise NotImplementedError("the reference instantiates SinusoidalEncoder(1, 0, 4, True) only")
        self.x_dim, self.min_deg, self.max_deg, self.use_identity = x_dim, min_deg, max_deg, use_identity
        # checkpoint compatibility (cednerf/encoder.py:18-20): the kernel has the octaves built in
        self.register_buffer("scales", torch.tensor([2 ** i for i in range(min_deg, max_deg)]))

    @property
    def latent_dim(self) -> int:
        return 9

    @torch.no_grad()
    def forward(self, x):
        return ops.time_embed(x)


class SinusoidalEncoderWithExp(SinusoidalEncoder):
    def __init__(self, x_dim=1, min_deg=0, max_deg=4, use_identity=True):
        super().__init__(x_dim, min_deg, max_deg, use_identity)
        self.register_buffer("scales_move", torch.tensor([i * 2 ** i for i in range(min_deg, max_deg)]))  # :58-60

    @torch.no_grad()
    def forward(self, x, move_norm):
        return ops.time_embed(x, move_norm)
