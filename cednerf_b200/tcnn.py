"""Drop-in for the subset of `tinycudann` the reference builds (cednerf/model.py:167-344): Encoding
(HashGrid / Frequency / SphericalHarmonics, optionally wrapped in a single-entry Composite), Network
(FullyFusedMLP, 64 neurons, ReLU, no output activation) and NetworkWithInputEncoding.

Same surface as tcnn's torch bindings: `.params` is one flat fp32 Parameter, `.n_input_dims`,
`.n_output_dims`, forward(x[N,in]) -> fp16 [N,out].  Everything runs in libcednerf_b200.so."""
from __future__ import annotations

import math

import torch

from . import ops


class Encoding(torch.nn.Module):
    def __init__(self, n_input_dims: int, encoding_config: dict, seed: int = 1337, dtype=torch.float16):
        super().__init__()
        self.n_input_dims = n_input_dims
        cfg = encoding_config
        if cfg["otype"] == "Composite":
            if len(cfg["nested"]) != 1:
                raise NotImplementedError("only single-entry Composite encodings (what the reference uses)")
            cfg = cfg["nested"][0]
        self.cfg = cfg
        ot = cfg["otype"]
        if ot == "HashGrid":
            self.n_levels = int(cfg["n_levels"])
            self.n_features = int(cfg.get("n_features_per_level", 2))
            if self.n_features != 2 or n_input_dims != 3:
                raise NotImplementedError("HashGrid: 3 input dims, 2 features per level")
            self.levels, total, self.level_info = ops.grid_levels(
                self.n_levels, cfg["base_resolution"], math.log(cfg["per_level_scale"]), 2 ** cfg["log2_hashmap_size"])
            g = torch.Generator().manual_seed(seed)
            self.params = torch.nn.Parameter((torch.rand(total * self.n_features, generator=g) * 2 - 1) * 1e-4)
            self.n_output_dims = self.n_levels * self.n_features
            self._f16 = ops._F16Cache()
            self.params._cednerf_f16 = self._f16  # lets optim.FusedAdam write the fp16 copy inside its update pass
        elif ot == "Frequency":
            self.n_frequencies = int(cfg["n_frequencies"])
            self.n_output_dims = n_input_dims * 2 * self.n_frequencies
            self.params = torch.nn.Parameter(torch.zeros(0))
        elif ot == "SphericalHarmonics":
            if cfg["degree"] != 2 or n_input_dims != 3:
                raise NotImplementedError("SphericalHarmonics: degree 2 on 3 dims (what the reference uses)")
            self.n_output_dims = 4
            self.params = torch.nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(ot)

    def table_f16(self) -> torch.Tensor:
        return self._f16.get(self.params, ops.cast_f16)

    def forward(self, x):
        ot = self.cfg["otype"]
        if ot == "HashGrid":
            return ops.HashGridFunction.apply(x, self.params.view(-1, 2), self.table_f16(), self.levels, False, False,
                                              False)
        if ot == "Frequency":
            return ops.FrequencyFunction.apply(x, self.n_frequencies, 0)
        return ops.sh2_encode(x)


class Network(torch.nn.Module):
    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: dict, seed: int = 1337):
        super().__init__()
        if network_config.get("activation", "ReLU") != "ReLU" or network_config.get("output_activation", "None") != "None":
            raise NotImplementedError("ReLU hidden / no output activation (what the reference uses)")
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.n_neurons, self.n_hidden = int(network_config["n_neurons"]), int(network_config["n_hidden_layers"])
        self.desc, n_params = ops.mlp_desc(n_input_dims, n_output_dims, self.n_neurons, self.n_hidden)
        g = torch.Generator().manual_seed(seed)
        chunks = []
        for l in range(self.desc.n_layers):  # Xavier-uniform on the padded [out, in] matrices (tcnn)
            o, i = self.desc.dim_out[l], self.desc.dim_in[l]
            chunks.append(((torch.rand(o, i, generator=g) * 2 - 1) * math.sqrt(6.0 / (o + i))).reshape(-1))
        self.params = torch.nn.Parameter(torch.cat(chunks))
        assert self.params.numel() == n_params
        self._image = ops._F16Cache()

    def weight_image(self) -> torch.Tensor:
        return self._image.get(self.params, lambda p: ops.mlp_pack(p, self.desc))

    def forward(self, x):
        save = torch.is_grad_enabled() and (x.requires_grad or self.params.requires_grad)
        return ops.MlpFunction.apply(x, self.params, self.weight_image(), self.desc, save, self.n_output_dims)


class NetworkWithInputEncoding(torch.nn.Module):
    """tinycudann.NetworkWithInputEncoding: ONE module with ONE flat `params` Parameter (encoding parameters, none for
    Frequency, followed by the network's), so a reference checkpoint's `xyz_wrap.params` / `mlp_feat_prediction.params`
    keys load with strict=True.  `encoding` and `network` are helpers that are deliberately NOT registered as
    sub-modules (they would add `*.network.params` / `*.encoding.params` keys tcnn does not have); `network` borrows
    this module's Parameter object, which `.to()` / `.cuda()` update in place."""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed: int = 1337):
        super().__init__()
        encoding = Encoding(n_input_dims, encoding_config, seed)
        if encoding.params.numel():
            raise NotImplementedError("only parameter-free input encodings (the reference uses Frequency)")
        network = Network(encoding.n_output_dims, n_output_dims, network_config, seed)
        self.params = network.params  # registered here, once
        del network._parameters["params"]
        network.__dict__["params"] = self.params
        self.__dict__["encoding"], self.__dict__["network"] = encoding, network
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        if self.network.__dict__["params"] is not self.params:  # parameter objects were replaced, not updated in place
            self.network.__dict__["params"] = self.params
        return out

    def forward(self, x):
        if self.encoding.cfg["otype"] == "Frequency":  # write the padded MLP operand directly
            x16 = ops.FrequencyFunction.apply(x, self.encoding.n_frequencies, self.network.desc.dim_in[0])
            return self.network(x16)
        return self.network(self.encoding(x))
