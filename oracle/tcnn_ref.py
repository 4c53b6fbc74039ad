"""ORACLE (test infrastructure) — restatement of the tiny-cuda-nn modules the reference uses.

tinycudann is an un-vendored, un-pinned dependency (git master, cednerf/model.py:15-23) and is
CUDA-only; its published algorithms are restated from SURVEY.md Appendix B.  PARITY UNPINNED.

Reference call sites (all in cednerf/model.py):
  tcnn.Encoding(HashGrid)                :242-252, used :384
  tcnn.Encoding(Composite[SH deg 2])     :226-239, used :450-455
  tcnn.NetworkWithInputEncoding(Freq)    :200-222 (xyz_wrap), :312-327, :329-344
  tcnn.Network (FullyFusedMLP 64, ReLU)  :280-290 (mlp_base), :292-309 (mlp_head)

Numerical contract shared with cednerf_b200 (documented in DESIGN.md):
  * hash grid: pos = x*scale + 0.5 as two rounded fp32 ops; trilinear weights fp32 in corner-bit
    order; table values fp16; accumulation fp32, ONE rounding to fp16 at the output (tcnn/Taichi
    accumulate in fp16; fp32 is strictly more accurate and inside the 2e-3 tolerance).
  * Frequency: sin(pi * (2^k x + phase/2)) evaluated in fp64 then rounded to fp16.
  * MLP: fp16 weights/activations, fp32 accumulate, ReLU hidden, no bias, inputs padded to a
    multiple of 16 with the constant 1.0 (tcnn behaviour), outputs padded to 16 and sliced;
    gradients between layers rounded to fp16, weight gradients fp32.
  * dtype flow of gradients, as in tcnn's torch bindings: every encoding / network OUTPUT is an fp16
    tensor, so the gradient that reaches it is fp16-rounded; the network's INPUT gradient is produced in
    the network precision (tcnn: dL_dinput is a __half matrix) and is fp16-rounded as well; the hash
    grid's and the Frequency encoding's input gradients (dL/dx) are fp32.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

PRIME_Y, PRIME_Z = 2654435761, 805459861


# --------------------------------------------------------------------------------------
# level geometry (E0)
# --------------------------------------------------------------------------------------
def grid_levels(n_levels: int, base_resolution: float, log_per_level_scale: float, max_params: int):
    """Per-level (scale f32, resolution, size, offset, hashed) — cednerf/taichi_kernel/hash_encoder_half.py:12-35,268-293."""
    scales, ress, sizes, offsets, hashed = [], [], [], [], []
    off = 0
    for l in range(n_levels):
        s = float(base_resolution) * math.exp(float(l) * log_per_level_scale) - 1.0
        res = int(math.ceil(s)) + 1
        full = res ** 3
        size = min(int(max_params), (full + 7) // 8 * 8)
        scales.append(np.float32(s))
        ress.append(res)
        sizes.append(size)
        offsets.append(off)
        hashed.append(full > size)
        off += size
    return scales, ress, sizes, offsets, hashed, off


class _RoundF16(torch.autograd.Function):
    """fp16 rounding in both directions (activations forward, gradients backward)."""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g.half().float()


class _RoundF16Fwd(torch.autograd.Function):
    """fp16 rounding forward only; the gradient passes through in fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _corner_index(gx, gy, gz, res, size, is_hashed):
    if is_hashed:
        h = (gx ^ (gy * PRIME_Y) ^ (gz * PRIME_Z)) & 0xFFFFFFFF
    else:
        h = (gx + gy * res + gz * res * res) & 0xFFFFFFFF
    return h % size


def hashgrid_forward(x: torch.Tensor, table: torch.Tensor, levels, n_features: int = 2,
                     cell_round_f16: bool = False, round_output: bool = True) -> torch.Tensor:
    """E1.  x f32[N,3] in [0,1]; table [sum sizes, F] (rounded to fp16 here).  Differentiable in both."""
    scales, ress, sizes, offsets, hashed, _ = levels
    tab = _RoundF16Fwd.apply(table.float())
    xf = x.float()
    outs = []
    for l in range(len(scales)):
        scale = float(scales[l])
        pos = xf * scale
        pos = pos + 0.5
        g = torch.floor(pos.detach())
        gsub = g.half().float() if cell_round_f16 else g  # E1q: Taichi rounds the cell to f16
        f = pos - gsub
        gi = g.to(torch.int64)
        acc = torch.zeros(x.shape[0], n_features, dtype=torch.float32, device=x.device)
        for c in range(8):
            w = torch.ones(x.shape[0], dtype=torch.float32, device=x.device)
            cg = []
            for d in range(3):
                if c & (1 << d):
                    w = w * f[:, d]
                    cg.append(gi[:, d] + 1)
                else:
                    w = w * (1.0 - f[:, d])
                    cg.append(gi[:, d])
            idx = _corner_index(cg[0], cg[1], cg[2], ress[l], sizes[l], hashed[l]) + offsets[l]
            acc = acc + w[:, None] * tab[idx]
        outs.append(acc)
    out = torch.cat(outs, -1)
    return _RoundF16.apply(out) if round_output else out


def frequency_encode(x: torch.Tensor, n_frequencies: int = 4) -> torch.Tensor:
    """E6.  out[j]: dim=j//(2n), k=(j//2)%n, phase=(j%2)*pi/2 -> sin(2^k*pi*x + phase).  f32 (fp16-rounded)."""
    n = n_frequencies
    xd = x.double()
    cols = []
    for j in range(x.shape[-1] * 2 * n):
        dim, k, p = j // (2 * n), (j // 2) % n, j % 2
        cols.append(torch.sin(math.pi * (xd[:, dim] * float(2 ** k) + 0.5 * p)))
    return _RoundF16.apply(torch.stack(cols, -1).float())


def sh_encode_deg2(d01: torch.Tensor) -> torch.Tensor:
    """E7.  Input in [0,1]; x,y,z = 2*in-1."""
    v = d01.float() * 2.0 - 1.0
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    out = torch.stack([torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y,
                       0.48860251190291987 * z, -0.48860251190291987 * x], -1)
    return _RoundF16.apply(out)


def pad16(n: int) -> int:
    return (n + 15) // 16 * 16


def mlp_layer_shapes(n_in: int, n_out: int, n_neurons: int, n_hidden: int):
    dims = [pad16(n_in)] + [n_neurons] * n_hidden + [pad16(n_out)]
    return [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def mlp_init_params(n_in, n_out, n_neurons=64, n_hidden=1, generator=None) -> torch.Tensor:
    """Xavier-uniform per layer on the padded [out,in] matrices, concatenated flat fp32."""
    chunks = []
    for (o, i) in mlp_layer_shapes(n_in, n_out, n_neurons, n_hidden):
        a = math.sqrt(6.0 / (o + i))
        chunks.append(((torch.rand(o, i, generator=generator) * 2 - 1) * a).reshape(-1))
    return torch.cat(chunks)


def mlp_forward(x: torch.Tensor, params: torch.Tensor, n_in, n_out, n_neurons=64, n_hidden=1,
                output_f16: bool = True) -> torch.Tensor:
    """F-table numerics.  x [N,n_in] (any float) -> [N,n_out] f32 holding fp16-rounded values."""
    shapes = mlp_layer_shapes(n_in, n_out, n_neurons, n_hidden)
    h = _RoundF16.apply(x.float())
    if shapes[0][1] > n_in:
        h = torch.cat([h, torch.ones(h.shape[0], shapes[0][1] - n_in, device=h.device)], -1)
    off = 0
    for li, (o, i) in enumerate(shapes):
        w = _RoundF16Fwd.apply(params[off:off + o * i].view(o, i).float())
        off += o * i
        h = h @ w.t()
        if li < len(shapes) - 1:
            h = _RoundF16.apply(torch.relu(h))
    out = h[:, :n_out]
    return _RoundF16.apply(out) if output_f16 else out


# --------------------------------------------------------------------------------------
# module shims with the tcnn constructor surface
# --------------------------------------------------------------------------------------
class Encoding(torch.nn.Module):
    def __init__(self, n_input_dims: int, encoding_config: dict, seed: int = 1337):
        super().__init__()
        self.n_input_dims = n_input_dims
        cfg = encoding_config
        if cfg["otype"] == "Composite":
            assert len(cfg["nested"]) == 1, "oracle supports single-entry Composite (as the reference uses)"
            cfg = cfg["nested"][0]
        self.cfg = cfg
        ot = cfg["otype"]
        if ot == "HashGrid":
            self.n_levels = cfg["n_levels"]
            self.n_features = cfg.get("n_features_per_level", 2)
            self.levels = grid_levels(self.n_levels, cfg["base_resolution"], math.log(cfg["per_level_scale"]),
                                      2 ** cfg["log2_hashmap_size"])
            g = torch.Generator().manual_seed(seed)
            self.params = torch.nn.Parameter((torch.rand(self.levels[5] * self.n_features, generator=g) * 2 - 1) * 1e-4)
            self.n_output_dims = self.n_levels * self.n_features
        elif ot == "Frequency":
            self.n_output_dims = n_input_dims * 2 * cfg["n_frequencies"]
            self.params = torch.nn.Parameter(torch.zeros(0))
        elif ot == "SphericalHarmonics":
            assert cfg["degree"] == 2 and n_input_dims == 3
            self.n_output_dims = 4
            self.params = torch.nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(ot)

    def forward(self, x):
        ot = self.cfg["otype"]
        if ot == "HashGrid":
            return hashgrid_forward(x, self.params.view(-1, self.n_features), self.levels, self.n_features)
        if ot == "Frequency":
            return frequency_encode(x, self.cfg["n_frequencies"])
        return sh_encode_deg2(x)


class Network(torch.nn.Module):
    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict, seed: int = 1337):
        super().__init__()
        assert network_config.get("activation", "ReLU") == "ReLU"
        assert network_config.get("output_activation", "None") == "None"
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.n_neurons, self.n_hidden = network_config["n_neurons"], network_config["n_hidden_layers"]
        g = torch.Generator().manual_seed(seed)
        self.params = torch.nn.Parameter(mlp_init_params(n_input_dims, n_output_dims, self.n_neurons, self.n_hidden, g))

    def forward(self, x):
        return mlp_forward(x, self.params, self.n_input_dims, self.n_output_dims, self.n_neurons, self.n_hidden)


class NetworkWithInputEncoding(torch.nn.Module):
    """One module, one flat `params` (tcnn's torch binding): checkpoint key `<name>.params`."""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed: int = 1337):
        super().__init__()
        encoding = Encoding(n_input_dims, encoding_config, seed)
        assert encoding.params.numel() == 0, "oracle: only parameter-free input encodings"
        network = Network(encoding.n_output_dims, n_output_dims, network_config, seed)
        self.params = network.params
        del network._parameters["params"]
        network.__dict__["params"] = self.params
        self.__dict__["encoding"], self.__dict__["network"] = encoding, network  # helpers, not sub-modules
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims

    def forward(self, x):
        self.network.__dict__["params"] = self.params
        return self.network(self.encoding(x))
