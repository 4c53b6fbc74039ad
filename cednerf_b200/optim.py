"""Optimiser side of the reference's training step (train_real.py:252, 267-274, 412-420): `grad_scaler.step(optimizer)`
with `apex.optimizers.FusedAdam(params, lr=lr, eps=1e-15)` (torch.optim.Adam when apex is missing), SURVEY.md §8f N2.

`FusedAdam` keeps apex's constructor surface; `GradScaler` is torch.amp.GradScaler with a cheaper `step` for it: one
read-only non-finite check and ONE pass that unscales, applies Adam to the fp32 master parameters and writes the fp16
working copy of the hash table the next forward reads.  Nothing is read back by the host (the step counter, the loss
scale and the found-inf flag stay on the device).  A stock torch GradScaler also works with `FusedAdam` (it sets
`grad_scale` / `found_inf` on the optimiser, as it does for torch's own fused Adam)."""
from __future__ import annotations

import ctypes

import torch
from torch.amp.grad_scaler import OptState

from . import _lib
from ._lib import AdamTensors, OPT_MAX_TENSORS, call, ptr, stream


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-8, adam_w_mode=True,
                 weight_decay=0.0, amsgrad=False, set_grad_none=True):
        if amsgrad:
            raise RuntimeError("FusedAdam does not support the AMSGrad variant.")  # apex's message
        if not bias_correction:
            raise NotImplementedError("bias_correction=False is not used by the reference")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.adam_w_mode, self.set_grad_none = bool(adam_w_mode), set_grad_none
        self._step_t = None

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step_t = None  # re-derived from the loaded "step" entries at the next step()

    def zero_grad(self, set_to_none=None):
        super().zero_grad(self.set_grad_none if set_to_none is None else set_to_none)

    def _tensors(self):
        """[(param, group)] with a gradient, checked for what the kernel handles."""
        out = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                if p.grad.is_sparse or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: dense fp32 parameters and gradients only")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FusedAdam: contiguous parameters and gradients only")
                out.append((p, group))
        return out

    def _table(self, items, with_state):
        """AdamTensors structs of at most OPT_MAX_TENSORS tensors each."""
        tables = []
        for i in range(0, len(items), OPT_MAX_TENSORS):
            t = AdamTensors()
            chunk = items[i:i + OPT_MAX_TENSORS]
            t.n_tensors = len(chunk)
            for k, (p, group) in enumerate(chunk):
                t.g[k], t.n[k] = ptr(p.grad), p.numel()
                if with_state:
                    st = self.state[p]
                    t.p[k], t.m[k], t.v[k] = ptr(p), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"])
                    t.lr[k], t.weight_decay[k] = float(group["lr"]), float(group["weight_decay"])
                    cache = getattr(p, "_cednerf_f16", None)
                    t.p16[k] = ptr(cache.val) if cache is not None else None
            tables.append(t)
        return tables

    @torch.no_grad()
    def check_nonfinite(self) -> torch.Tensor:
        """float32[1] on the device: 1 when any gradient element is inf / nan (GradScaler's check, read-only)."""
        items = self._tensors()
        dev = items[0][0].device if items else torch.device("cuda")
        found = torch.zeros(1, dtype=torch.float32, device=dev)
        for t in self._table(items, False):
            call("cednerf_nonfinite_check", ctypes.byref(t), ptr(found), stream())
        return found

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = self._tensors()
        if not items:
            return loss
        dev = items[0][0].device
        if self._step_t is None:
            # one shared device counter (every parameter steps together); after load_state_dict it resumes from the
            # loaded per-parameter "step" entries, which all hold the same value
            loaded = [st["step"] for st in self.state.values() if "step" in st]
            start = float(torch.as_tensor(loaded[0]).reshape(-1)[0]) if loaded else 0.0
            self._step_t = torch.full((1,), start, dtype=torch.float32, device=dev)
        for p, _ in items:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)
            st["step"] = self._step_t
            cache = getattr(p, "_cednerf_f16", None)  # hash table: fp16 working copy owned by the encoder
            if cache is not None:  # the update pass (re)writes every element of it, also when the step is skipped
                if cache.val is None or cache.val.numel() != p.numel() or cache.val.device != p.device:
                    cache.val = torch.empty(p.numel(), dtype=torch.float16, device=p.device)
        grad_scale, found_inf = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        betas, eps = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        for g in self.param_groups:
            if g["betas"] != betas or g["eps"] != eps:
                raise NotImplementedError("FusedAdam: one (betas, eps) for all groups")
        for i, t in enumerate(self._table(items, True)):
            call("cednerf_adam_step", ctypes.byref(t), ptr(self._step_t), int(i == 0),
                 ptr(grad_scale) if grad_scale is not None else None, ptr(found_inf) if found_inf is not None else None,
                 float(betas[0]), float(betas[1]), float(eps), int(self.adam_w_mode), stream())
        for p, _ in items:  # the kernel wrote through raw pointers: tell autograd / the version-keyed caches
            torch.autograd.graph.increment_version(p)
            cache = getattr(p, "_cednerf_f16", None)
            if cache is not None:
                cache.adopt(p)
        return loss


class GradScaler(torch.amp.GradScaler):
    """torch.cuda.amp.GradScaler(2**10) of train_real.py:252; `step(FusedAdam)` skips torch's read-modify-write inf
    check in favour of the optimiser's read-only one and hands it the scale, so unscale happens inside the update."""

    def __init__(self, init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enabled=True):
        super().__init__("cuda", init_scale, growth_factor, backoff_factor, growth_interval, enabled)

    def step(self, optimizer, *args, **kwargs):
        if not self._enabled or not isinstance(optimizer, FusedAdam) or "closure" in kwargs:
            return super().step(optimizer, *args, **kwargs)
        self._check_scale_growth_tracker("step")
        state = self._per_optimizer_states[id(optimizer)]
        if state["stage"] is OptState.STEPPED:
            raise RuntimeError("step() has already been called since the last update().")
        if state["stage"] is OptState.READY:
            found_inf = optimizer.check_nonfinite()
            state["found_inf_per_device"] = {found_inf.device: found_inf}
            optimizer.grad_scale = self._get_scale_async()
        else:  # unscale_() was called: gradients are already divided by the scale
            found_inf = sum(t.to(self._scale.device, non_blocking=True) for t in state["found_inf_per_device"].values())
            optimizer.grad_scale = None
        optimizer.found_inf = found_inf
        try:
            ret = optimizer.step(*args, **kwargs)
        finally:
            del optimizer.grad_scale
            del optimizer.found_inf
        state["stage"] = OptState.STEPPED
        return ret
