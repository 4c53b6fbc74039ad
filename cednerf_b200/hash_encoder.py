"""The reference's in-tree Taichi encoders as modules (same constructor surface):
   HashEncoder    cednerf/taichi_kernel/hash_encoder_half.py:231-385   (3-D, 2 features/level)
   HashEncoder4D  cednerf/taichi_kernel/hash_encoder_inter.py:279-430  (xyz+t, 4 key-frames x 2 features)
Default numerics are the tcnn-equivalent ones (fp32 cell fraction, correct input gradient); the 4-D encoder's
`taichi_compat` flag reproduces the t == 1 key-frame choice of the Taichi source."""
from __future__ import annotations

import math

import torch

from . import ops


class _Base(torch.nn.Module):
    def __init__(self, max_params, levels, base_res, max_res, feature_per_level, entry_width, seed):
        super().__init__()
        if feature_per_level != 2:
            raise NotImplementedError("2 features per level")
        self.log_b = math.log(float(max_res) / float(base_res)) / float(int(levels) - 1)
        self.levels_desc, total, self.level_info = ops.grid_levels(int(levels), float(base_res), self.log_b,
                                                                   int(max_params))
        self.hash_level, self.feature_per_level = int(levels), int(feature_per_level)
        self.out_dim = self.n_output_dims = 2 * int(levels)
        g = torch.Generator().manual_seed(seed)
        self.hash_table = torch.nn.Parameter((torch.rand(total, entry_width, generator=g) * 2 - 1) * 1e-4)
        self.register_buffer("offsets", torch.tensor([i[3] for i in self.level_info], dtype=torch.int32), persistent=False)
        self.register_buffer("hash_map_sizes", torch.tensor([i[2] for i in self.level_info], dtype=torch.int32),
                             persistent=False)
        self.begin_fast_hash_level = next((i for i, v in enumerate(self.level_info) if v[4]), int(levels))
        self._f16 = ops._F16Cache()
        self.hash_table._cednerf_f16 = self._f16  # optim.FusedAdam writes the fp16 copy inside its update pass

    def table_f16(self):
        return self._f16.get(self.hash_table, ops.cast_f16)


class HashEncoder(_Base):
    def __init__(self, max_params=2 ** 19, levels=16, base_res=16.0, max_res=2048.0, feature_per_level=2, seed=1337):
        super().__init__(max_params, levels, base_res, max_res, feature_per_level, 2, seed)

    def forward(self, positions):
        return ops.HashGridFunction.apply(positions, self.hash_table, self.table_f16(), self.levels_desc, False, False,
                                          False)


class HashEncoder4D(_Base):
    def __init__(self, max_params=2 ** 19, levels=16, base_res=16.0, max_res=2048.0, feature_per_level=2,
                 taichi_compat=False, seed=1337):
        super().__init__(max_params, levels, base_res, max_res, feature_per_level, 8, seed)
        self.taichi_compat = bool(taichi_compat)

    def forward(self, positions):
        return ops.HashGridFunction.apply(positions, self.hash_table, self.table_f16(), self.levels_desc, True,
                                          self.taichi_compat, False)
