"""ORACLE (test infrastructure) — restatement of the in-tree Taichi hash encoders.

Follows cednerf/taichi_kernel/hash_encoder_half.py (3-D, f16) and
cednerf/taichi_kernel/hash_encoder_inter.py (4-D xyz+t, 4 temporal key-frames per entry).
The Taichi compiler is not installable here, so the kernels cannot be executed:
PARITY UNPINNED; the source is in-tree and is followed line by line below.

Documented divergences (default = mathematically correct / tcnn-equivalent; `taichi_compat=True`
reproduces the in-tree quirk where it is well defined):
  E1q  hash_encoder_half.py:133  fraction uses the cell index rounded to f16      -> compat flag
  E1   hash_encoder_half.py:159  per-corner accumulation in f16                    -> compat flag
  E3   hash_encoder_inter.py:151-160  t==1 selects key-frame 2 with tau=0          -> compat flag
  E2q  hash_encoder_half.py:219-220, 367  backward w/(+-f) (0/0 at faces), missing scale,
       accumulation into torch.empty_like                                          -> NOT reproduced
"""
from __future__ import annotations

import math

import torch

from .tcnn_ref import _RoundF16, _RoundF16Fwd, _corner_index, grid_levels, hashgrid_forward


def encoder_levels(max_params=2 ** 19, levels=16, base_res=16.0, max_res=2048.0):
    """hash_encoder_half.py:24-35 (log scale) + :268-293 (sizes/offsets)."""
    log_b = math.log(float(max_res) / float(base_res)) / float(levels - 1)
    return grid_levels(int(levels), float(base_res), log_b, int(max_params)), log_b


class HashEncoder(torch.nn.Module):
    """3-D encoder, hash_encoder_half.py:231-385."""

    def __init__(self, max_params=2 ** 19, levels=16, base_res=16.0, max_res=2048.0, feature_per_level=2,
                 taichi_compat=False, seed=1337):
        super().__init__()
        self.levels_geom, self.log_b = encoder_levels(max_params, levels, base_res, max_res)
        self.hash_level, self.feature_per_level = int(levels), feature_per_level
        self.out_dim = self.n_output_dims = feature_per_level * int(levels)
        self.taichi_compat = taichi_compat
        g = torch.Generator().manual_seed(seed)
        total = self.levels_geom[5]
        self.hash_table = torch.nn.Parameter((torch.rand(total, feature_per_level, generator=g) * 2 - 1) * 1e-4)
        self.register_buffer("offsets", torch.tensor(self.levels_geom[3], dtype=torch.int32), persistent=False)
        self.register_buffer("hash_map_sizes", torch.tensor(self.levels_geom[2], dtype=torch.int32), persistent=False)
        self.begin_fast_hash_level = next((i for i, h in enumerate(self.levels_geom[4]) if h), int(levels))

    def forward(self, positions):
        if not self.taichi_compat:
            return hashgrid_forward(positions, self.hash_table, self.levels_geom, self.feature_per_level)
        return _forward_f16_accumulate(positions, self.hash_table, self.levels_geom, self.feature_per_level)


def _forward_f16_accumulate(x, table, levels, n_features):
    scales, ress, sizes, offsets, hashed, _ = levels
    tab = table.half()
    outs = []
    for l in range(len(scales)):
        pos = x.float() * float(scales[l]) + 0.5
        g = torch.floor(pos)
        f = pos - g.half().float()
        gi = g.to(torch.int64)
        acc = torch.zeros(x.shape[0], n_features, dtype=torch.float16)
        for c in range(8):
            w = torch.ones(x.shape[0])
            cg = []
            for d in range(3):
                w = w * (f[:, d] if c & (1 << d) else 1.0 - f[:, d])
                cg.append(gi[:, d] + (1 if c & (1 << d) else 0))
            idx = _corner_index(cg[0], cg[1], cg[2], ress[l], sizes[l], hashed[l]) + offsets[l]
            acc = acc + (w[:, None] * tab[idx].float()).half()
        outs.append(acc)
    return torch.cat(outs, -1).float()


def hashgrid4d_forward(xyzt, table, levels, taichi_compat=False):
    """E3: hash_encoder_inter.py:121-199.  table [sum sizes, 8] = 4 key-frames x 2 features."""
    scales, ress, sizes, offsets, hashed, _ = levels
    tab = _RoundF16Fwd.apply(table.float()).view(-1, 4, 2)
    x, t = xyzt[:, :3].detach().float(), xyzt[:, 3].detach().float()
    ts = t * 3.0
    k = torch.floor(ts)
    if taichi_compat:
        tau = ts - k
        k = torch.clamp(k, max=2.0)
    else:
        k = torch.clamp(k, max=2.0)
        tau = ts - k
    k = k.to(torch.int64)
    rows = torch.arange(x.shape[0])
    outs = []
    for l in range(len(scales)):
        pos = x * float(scales[l])
        pos = pos + 0.5
        g = torch.floor(pos)
        f = pos - g
        gi = g.to(torch.int64)
        acc = torch.zeros(x.shape[0], 2)
        for c in range(8):
            w = torch.ones(x.shape[0])
            cg = []
            for d in range(3):
                w = w * (f[:, d] if c & (1 << d) else 1.0 - f[:, d])
                cg.append(gi[:, d] + (1 if c & (1 << d) else 0))
            idx = _corner_index(cg[0], cg[1], cg[2], ress[l], sizes[l], hashed[l]) + offsets[l]
            lo, hi = tab[idx, k], tab[idx, k + 1]
            acc = acc + w[:, None] * (lo * (1.0 - tau)[:, None] + hi * tau[:, None])
        outs.append(acc)
    return _RoundF16.apply(torch.cat(outs, -1))


class HashEncoder4D(torch.nn.Module):
    """4-D key-frame encoder, hash_encoder_inter.py:279-430 (class HashEncoder there)."""

    def __init__(self, max_params=2 ** 19, levels=16, base_res=16.0, max_res=2048.0, feature_per_level=2,
                 taichi_compat=False, seed=1337):
        super().__init__()
        assert feature_per_level == 2
        self.levels_geom, self.log_b = encoder_levels(max_params, levels, base_res, max_res)
        self.hash_level, self.taichi_compat = int(levels), taichi_compat
        self.out_dim = self.n_output_dims = 2 * int(levels)
        g = torch.Generator().manual_seed(seed)
        self.hash_table = torch.nn.Parameter((torch.rand(self.levels_geom[5], 8, generator=g) * 2 - 1) * 1e-4)

    def forward(self, positions):
        return hashgrid4d_forward(positions, self.hash_table, self.levels_geom, self.taichi_compat)
