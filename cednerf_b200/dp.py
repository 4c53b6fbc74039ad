"""One-process-per-GPU plumbing for the two ways the path shards.

* render: frames / row blocks are independent units - `shard_range` / `shard_interleaved`, no collective;
* train: pure data parallelism over rays - every rank runs the whole path on its slice of the batch and the
  gradients (hash table 191 MB fp32 + ~0.1 MB of MLP weights) are summed with one NCCL all-reduce per parameter,
  launched from a post-accumulate hook the moment that parameter's gradient is complete, so the big hash-table
  all-reduce overlaps the deformation-MLP backward that follows the encoder in the backward order.
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

    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, average: bool = True):
        self.world, self.average = world_size, average
        self.params = [p for p in params if p.requires_grad and p.numel() > 0]
        self.handles = []
        self._hooks = []
        if world_size > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._launch))

    def _launch(self, p: torch.nn.Parameter):
        if self.average:
            p.grad.div_(self.world)
        self.handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        """Call after backward(), before the optimiser (and before GradScaler's inf check)."""
        for h in self.handles:
            h.wait()
        self.handles.clear()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


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
