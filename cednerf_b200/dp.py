"""One-process-per-GPU plumbing for the two ways the path shards.

* render: frames / row blocks are independent units - `shard_range` / `shard_interleaved`, no collective;
* train: pure data parallelism over rays - every rank runs the whole path on its slice of the batch and the
  gradients (hash table 191 MB fp32 + ~0.1 MB of MLP weights) are summed with one NCCL all-reduce per parameter.
  The fused training backward produces all of them inside one autograd node, so the hash-table all-reduce is started
  from INSIDE it (ops.set_table_grad_hook): right after the table-gradient kernel, before the encoding's dL/dx and the
  deformation-net backward are launched, which then run under the transfer.  The small MLP gradients follow from
  post-accumulate hooks.
The reference is single-GPU; there is no reference collective to mirror."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n units for `rank`; block sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_interleaved(n: int, rank: int, world: int) -> List[int]:
    """Units rank, rank+world, ... (frames of a video: neighbouring poses cost alike, so interleaving balances)."""
    return list(range(rank, n, world))


class GradAllReducer:
    """Sum-all-reduce of parameter gradients, overlapped with the rest of the backward pass."""

    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, average: bool = True,
                 early_table: bool = True):
        self.world, self.average = world_size, average
        self.params = [p for p in params if p.requires_grad and p.numel() > 0]
        self.handles = []
        self._hooks = []
        self._early = None  # (table gradient tensor, its all-reduce) started from inside the fused backward
        if world_size > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._launch))
            if early_table:
                from . import ops

                ops.set_table_grad_hook(self._launch_table)
                self._ops = ops

    def _reduce(self, g: torch.Tensor):
        if self.average:
            g.div_(self.world)
        return dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)

    def _launch_table(self, g_table: torch.Tensor):
        owner = next((p for p in self.params if p.numel() == g_table.numel()), None)
        if owner is None or owner.grad is not None:
            return  # gradient accumulation: autograd adds into the existing .grad, which is reduced afterwards as a whole
        self._early = (g_table, self._reduce(g_table))

    def _launch(self, p: torch.nn.Parameter):
        if self._early is not None and p.grad.numel() == self._early[0].numel():
            g, work = self._early
            self._early = None
            work.wait()  # stream-level: the kernels of the rest of the backward are already enqueued ahead of this
            if p.grad.data_ptr() != g.data_ptr():
                p.grad.copy_(g)  # autograd kept a copy of the (then unreduced) gradient instead of the tensor itself
            return
        self.handles.append(self._reduce(p.grad))

    def wait(self):
        """Call after backward(), before the optimiser (and before GradScaler's inf check)."""
        for h in self.handles:
            h.wait()
        self.handles.clear()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
        if getattr(self, "_ops", None) is not None:
            self._ops.set_table_grad_hook(None)
            self._ops = None


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
