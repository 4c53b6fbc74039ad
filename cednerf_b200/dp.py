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

import math
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


# ----------------------------------------------------------------------------------------------------------------------
# NVLink peer memory + the fused reduce-scatter / Adam / all-gather step (csrc/dp.cu)
# ----------------------------------------------------------------------------------------------------------------------
class _RawCuda:
    """A raw device range as a __cuda_array_interface__ object (torch.as_tensor keeps it, and through it `owner`, alive)."""

    def __init__(self, address: int, nbytes: int, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 3,
                                         "strides": None}


class PeerMemory:
    """One buffer per rank, split into named regions with the same layout on every rank, mapped into every other rank.

    Two providers, tried in this order: (1) torch symmetric memory (`torch.distributed._symmetric_memory`): besides the
    peers' unicast mappings it gives a MULTICAST mapping of the buffers, which lets kernels reduce and broadcast inside
    the NVSwitch (NVLS: `multimem.ld_reduce` / `multimem.st`); (2) `cednerf_peer_alloc` + CUDA IPC (handles exchanged
    with all_gather_object): unicast peer mappings only.  PyTorch is the memory / rendezvous plumbing in (1); every
    byte that moves between GPUs is moved by this library's kernels."""

    ALIGN = 256

    def __init__(self, regions, device, group=None, provider: str = "auto"):
        """regions: [(name, nbytes)] - identical on every rank."""
        from . import _lib

        self._lib, self.device, self.group = _lib, torch.device(device), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.DP_MAX_RANKS:
            raise RuntimeError(f"peer-memory data parallelism supports up to {_lib.DP_MAX_RANKS} ranks")
        self.offsets, off = {}, 0
        for name, nbytes in regions:
            self.offsets[name] = (off, int(nbytes))
            off += (int(nbytes) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.nbytes = max(off, self.ALIGN)
        self._regions = [list(r) for r in regions]
        self._opened, self.mc_base, self.provider = [], 0, None
        ok = torch.zeros(1, device=self.device)
        if provider in ("auto", "symm"):
            try:
                self._init_symm()
                ok.fill_(1)
            except Exception as e:  # noqa: BLE001 - provider missing on this build / box: use IPC
                self._symm_error = repr(e)[:200]
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # every rank must end up on the same provider
            if float(ok.item()) == 1.0:
                self.provider = "symm"
        if self.provider is None:
            if provider == "symm":
                raise RuntimeError(f"symmetric memory unavailable: {getattr(self, '_symm_error', 'a peer failed')}")
            self._init_ipc()
            self.provider = "ipc"

    def _init_symm(self):
        import torch.distributed._symmetric_memory as symm

        grp = self.group if self.group is not None else dist.group.WORLD
        with torch.cuda.device(self.device):
            t = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            h = symm.rendezvous(t, grp)
            t.zero_()
            torch.cuda.synchronize()
            h.barrier()
        if h.world_size != self.world or len(h.buffer_ptrs) != self.world:
            raise RuntimeError("symmetric-memory group does not match the process group")
        self._raw, self._handle = t, h
        self.base = int(t.data_ptr())
        self.bases = [int(p) for p in h.buffer_ptrs]
        self.mc_base = int(h.multicast_ptr or 0)

    def _init_ipc(self):
        import ctypes

        _lib = self._lib
        with torch.cuda.device(self.device):
            base = ctypes.c_void_p()
            _lib.call("cednerf_peer_alloc", self.nbytes, ctypes.byref(base))
            self.base = int(base.value)
            handle = (ctypes.c_ubyte * 64)()
            _lib.call("cednerf_ipc_export", self.base, handle)
            mine = (bytes(handle), self.nbytes, self._regions)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=self.group)
            self.bases = [0] * self.world
            for r, (h, nb, regs) in enumerate(everyone):
                if nb != self.nbytes or regs != mine[2]:
                    raise RuntimeError(f"rank {r} laid out its peer buffer differently ({nb} vs {self.nbytes} bytes)")
                if r == self.rank:
                    self.bases[r] = self.base
                    continue
                mapped = ctypes.c_void_p()
                _lib.call("cednerf_ipc_open", (ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(mapped))
                self.bases[r] = int(mapped.value)
                self._opened.append(int(mapped.value))
        self._raw = torch.as_tensor(_RawCuda(self.base, self.nbytes, self), device=self.device)  # uint8 view of it all

    def local(self, name: str, dtype) -> torch.Tensor:
        """This rank's region as a flat tensor of `dtype` (zero-copy)."""
        off, nbytes = self.offsets[name]
        return self._raw[off:off + nbytes].view(dtype)

    def address(self, rank: int, name: str) -> int:
        """Where rank `rank`'s region is mapped in THIS process."""
        return self.bases[rank] + self.offsets[name][0]

    def multicast_address(self, name: str) -> int:
        """The region's address in the multicast mapping (loads reduce over / stores reach every rank), 0 if none."""
        return self.mc_base + self.offsets[name][0] if self.mc_base else 0

    def close(self):
        for b in self._opened:
            self._lib.call("cednerf_ipc_close", b)
        self._opened = []


from .optim import FusedAdam as _FusedAdam  # noqa: E402


class DistributedFusedAdam(_FusedAdam):
    """FusedAdam for one-process-per-GPU data parallelism: `GradScaler.step(opt)` / `opt.step()` do, stream-ordered and
    without any NCCL call, barrier -> fused [reduce-scatter of the table gradient over NVLink + unscale + Adam on the
    owned 1/N of the table + all-gather of the new fp32 / fp16 values into every replica] -> barrier (csrc/dp.cu).  The
    small MLP parameters are summed from all peers by every rank (fixed rank order: bit-identical replicas) and updated
    locally.  The Adam moments of the table exist only for the owned range (`gather_state()` reassembles them).

    The table parameter is the one carrying an fp16 working copy (`_cednerf_f16`, set by tcnn.Encoding / HashEncoder);
    its storage, its fp16 copy and its gradient are (re)homed in peer-visible memory on the first step."""

    def __init__(self, params, *args, group=None, average: bool = True, timeout_ms: int = 20000,
                 replicate_master: bool = False, provider: str = "auto", nvls="auto", **kwargs):
        """replicate_master=False (default): only the fp16 working copy - what forward and backward read - is all-gathered
        every step (2 B per element and peer instead of 6); a rank's fp32 master is then current for its OWNED range only,
        and `sync_master()` (one-sided, no collective) brings the rest up to date before a checkpoint is written.
        replicate_master=True stores the fp32 values into every replica each step as well."""
        super().__init__(params, *args, **kwargs)
        self.group, self.average, self.timeout_ms = group, bool(average), int(timeout_ms)
        self.replicate_master = bool(replicate_master)
        self.provider, self.nvls = provider, nvls   # provider: "auto" | "symm" | "ipc" (see PeerMemory)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._peer = None
        self._epoch = 0
        self._checked = False

    # ---- setup --------------------------------------------------------------------------------------------------
    def _all_params(self):
        return [p for g in self.param_groups for p in g["params"] if p.numel() > 0 and p.requires_grad]

    def setup(self):
        """Collective: allocate and exchange the peer buffers (call once, after the parameters are on the device)."""
        if self._peer is not None or self.world == 1:
            return
        import ctypes

        from . import _lib, ops

        params = self._all_params()
        tables = [p for p in params if getattr(p, "_cednerf_f16", None) is not None]
        if len(tables) != 1:
            raise RuntimeError("DistributedFusedAdam needs exactly one hash-table parameter (with an fp16 working copy)")
        self._table = tables[0]
        self._small = [p for p in params if p is not self._table]
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("DistributedFusedAdam: fp32 CUDA parameters only")
        n = self._table.numel()
        self._small_off, off = [], 0
        for p in self._small:
            self._small_off.append(off)
            off += (p.numel() + 3) // 4 * 4
        self._small_n = max(off, 4)
        dev = self._table.device
        regions = [("ctrl", int(_lib.load().cednerf_dp_ctrl_bytes())), ("grad", 4 * n), ("small", 4 * self._small_n),
                   ("p32", 4 * n), ("p16", 2 * n)]
        self._peer = PeerMemory(regions, dev, self.group, self.provider)
        pm = self._peer
        self._g_table, self._g_small = pm.local("grad", torch.float32), pm.local("small", torch.float32)
        self._p32, self._p16 = pm.local("p32", torch.float32), pm.local("p16", torch.float16)
        self._ctrl_f32 = pm.local("ctrl", torch.float32)          # arrive[8] | found_inf | timed_out
        self._ctrl_i32 = pm.local("ctrl", torch.int32)
        self._found = torch.zeros(1, dtype=torch.float32, device=dev)
        self._rehome()
        # owned ranges: multiples of 4 elements
        q = (n // 4) // self.world
        self._ranges = [(4 * q * r, n if r == self.world - 1 else 4 * q * (r + 1)) for r in range(self.world)]
        self._lo, self._hi = self._ranges[self.rank]
        # the fp16 copy is maintained by this optimiser; a backward pass without a step must not trigger a re-cast from a
        # master whose non-owned ranges may be one step behind (see ops._F16Cache)
        self._table._cednerf_f16.version_only = True
        peers = _lib.DpPeers()
        peers.world, peers.rank = self.world, self.rank
        for r in range(self.world):
            peers.ctrl[r] = pm.address(r, "ctrl")
        self._peers = peers
        ops.set_table_grad_alloc(self._alloc_table_grad)
        self._ops = ops
        self.barrier()  # every rank has mapped every buffer before anyone writes into a peer

    def _rehome(self):
        """Make the table parameter and its fp16 copy live in the peer buffers (again, if someone replaced them)."""
        from . import ops

        t = self._table
        if t.data_ptr() != self._p32.data_ptr():
            with torch.no_grad():
                self._p32.copy_(t.detach().reshape(-1))
                t.data = self._p32.view(t.shape)
        cache = t._cednerf_f16
        if cache.val is None or cache.val.data_ptr() != self._p16.data_ptr():
            fresh = cache.get(t, ops.cast_f16)  # current contents, cast by the library
            self._p16.copy_(fresh.reshape(-1))
            cache.val = self._p16
            cache.adopt(t)

    def _alloc_table_grad(self, shape, device):
        """Called by the fused training backward: the zeroed gradient buffer the peers can read."""
        if self._peer is None or math.prod(shape) != self._g_table.numel() or torch.device(device) != self._g_table.device:
            return None
        self._g_table.zero_()
        return self._g_table.view(shape)  # a fresh view object: autograd may adopt it as .grad without copying

    # ---- the step -----------------------------------------------------------------------------------------------
    def barrier(self):
        from ._lib import call, stream

        self._epoch += 1
        call("cednerf_dp_barrier", ctypes_byref(self._peers), self._epoch & 0xFFFFFFFF, self.timeout_ms, stream())

    def _stage(self):
        """This rank's gradients -> its peer-visible buffers (no copy when autograd adopted the buffer itself)."""
        with torch.no_grad():
            g = self._table.grad
            if g is None:
                self._g_table.zero_()
            elif g.data_ptr() != self._g_table.data_ptr():
                self._g_table.copy_(g.reshape(-1))
            for p, off in zip(self._small, self._small_off):
                dst = self._g_small[off:off + p.numel()]
                if p.grad is None:
                    dst.zero_()
                else:
                    dst.copy_(p.grad.reshape(-1))

    def _begin(self, with_check: bool):
        from ._lib import AdamTensors, call, ptr, stream

        self.setup()
        self._rehome()
        self._stage()
        dev = self._table.device
        if self._step_t is None:
            loaded = [st["step"] for st in self.state.values() if "step" in st]
            start = float(torch.as_tensor(loaded[0]).reshape(-1)[0]) if loaded else 0.0
            self._step_t = torch.full((1,), start, dtype=torch.float32, device=dev)
        found_local = self._ctrl_f32[8:9]
        found_local.zero_()
        if with_check:
            t = AdamTensors()
            t.n_tensors = 2
            t.g[0], t.n[0] = ptr(self._g_table), self._g_table.numel()
            t.g[1], t.n[1] = ptr(self._g_small), self._g_small.numel()
            call("cednerf_nonfinite_check", ctypes_byref(t), ptr(found_local), stream())
        self.barrier()   # every rank's gradients (and its flag) are complete and visible
        call("cednerf_dp_found_inf", ctypes_byref(self._peers), ptr(self._found), ptr(self._step_t), stream())
        self._checked = True

    @torch.no_grad()
    def check_nonfinite(self) -> torch.Tensor:
        """GradScaler's check, made global: 1 when ANY rank's gradient holds inf / nan (every rank gets the same answer)."""
        if self.world == 1:
            return super().check_nonfinite()
        self._begin(True)
        return self._found

    @torch.no_grad()
    def step(self, closure=None):
        if self.world == 1:
            return super().step(closure)
        from ._lib import DpAdam, call, ptr, stream

        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self._checked:
            self._begin(False)
        self._checked = False
        grad_scale, found_inf = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        found_inf = self._found if found_inf is None else found_inf
        betas, eps = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        pm, dev = self._peer, self._table.device
        group_of = {id(p): g for g in self.param_groups for p in g["params"]}

        def moments(p, n):
            st = self.state[p]
            if "exp_avg" not in st or st["exp_avg"].numel() != n:
                st["exp_avg"] = torch.zeros(n, dtype=torch.float32, device=dev)
                st["exp_avg_sq"] = torch.zeros(n, dtype=torch.float32, device=dev)
            st["step"] = self._step_t
            return st["exp_avg"], st["exp_avg_sq"]

        def launch(p, grad_name, grad_off, lo, hi, broadcast):
            g = group_of[id(p)]
            if g["betas"] != betas or g["eps"] != eps:
                raise NotImplementedError("FusedAdam: one (betas, eps) for all groups")
            m, v = moments(p, hi - lo)
            a = DpAdam()
            a.world, a.rank = self.world, self.rank
            for r in range(self.world):  # rank order: the small tensors are summed by every rank, bit-identically
                a.grad[r] = pm.address(r, grad_name) + 4 * grad_off
            if broadcast:  # entry 0 = the local replica, then the peers
                order = [self.rank] + [r for r in range(self.world) if r != self.rank]
                a.n_out = self.world
                for k, r in enumerate(order):
                    a.p32_out[k] = pm.address(r, "p32") if (k == 0 or self.replicate_master) else None
                    a.p16_out[k] = pm.address(r, "p16")
            else:
                a.n_out = 1
                a.p32_out[0], a.p16_out[0] = p.data_ptr(), None
            a.m, a.v, a.lo, a.hi = ptr(m), ptr(v), lo, hi
            a.lr, a.weight_decay = float(g["lr"]), float(g["weight_decay"])
            a.grad_div = float(self.world) if self.average else 1.0
            if broadcast and self._use_nvls():   # in-switch reduce / broadcast
                a.grad_mc = pm.multicast_address(grad_name) + 4 * grad_off
                a.p16_mc = pm.multicast_address("p16")
            call("cednerf_dp_adam", ctypes_byref(a), ptr(self._step_t), ptr(grad_scale) if grad_scale is not None else None,
                 ptr(found_inf), float(betas[0]), float(betas[1]), float(eps), int(self.adam_w_mode), stream())

        launch(self._table, "grad", 0, self._lo, self._hi, True)
        from ._lib import DpSmall

        for i0 in range(0, len(self._small), 8):   # the small tensors: one launch per 8
            sm = DpSmall()
            chunk = list(zip(self._small, self._small_off))[i0:i0 + 8]
            sm.world, sm.n_tensors = self.world, len(chunk)
            for r in range(self.world):
                sm.grad[r] = pm.address(r, "small")
            for k, (p, off) in enumerate(chunk):
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam: contiguous parameters only")
                g = group_of[id(p)]
                if g["betas"] != betas or g["eps"] != eps:
                    raise NotImplementedError("FusedAdam: one (betas, eps) for all groups")
                m, v = moments(p, p.numel())
                sm.p[k], sm.m[k], sm.v[k], sm.off[k], sm.n[k] = p.data_ptr(), ptr(m), ptr(v), off, p.numel()
                sm.lr[k], sm.weight_decay[k] = float(g["lr"]), float(g["weight_decay"])
            sm.grad_div = float(self.world) if self.average else 1.0
            call("cednerf_dp_adam_small", ctypes_byref(sm), ptr(self._step_t), ptr(grad_scale) if grad_scale is not None else None,
                 ptr(found_inf), float(betas[0]), float(betas[1]), float(eps), int(self.adam_w_mode), stream())
        self.barrier()   # every replica holds every owner's update before anyone's next forward reads it
        for p in [self._table] + self._small:
            torch.autograd.graph.increment_version(p)
        self._table._cednerf_f16.adopt(self._table)
        return loss

    # ---- housekeeping -------------------------------------------------------------------------------------------
    def gather_state(self):
        """Collective: the table's full Adam moments (exp_avg, exp_avg_sq) reassembled from the ranks' owned ranges."""
        st = self.state[self._table]
        n = self._table.numel()
        out = []
        for key in ("exp_avg", "exp_avg_sq"):
            full = torch.zeros(n, dtype=torch.float32, device=self._table.device)
            full[self._lo:self._hi] = st[key]
            dist.all_reduce(full, group=self.group)
            out.append(full)
        return tuple(out)

    @torch.no_grad()
    def sync_master(self):
        """One-sided (no collective, safe between steps: a peer cannot enter its next update before this rank reaches the
        barrier in front of it): pull every peer's OWNED range of the fp32 master into this rank's copy, after which
        `state_dict()` of the model is complete here.  A no-op with replicate_master=True."""
        if self._peer is None or self.replicate_master:
            return
        for r, (lo, hi) in enumerate(self._ranges):
            if r == self.rank or hi == lo:
                continue
            src = torch.as_tensor(_RawCuda(self._peer.address(r, "p32") + 4 * lo, 4 * (hi - lo), self._peer),
                                  device=self._table.device).view(torch.float32)
            self._p32[lo:hi].copy_(src)
        torch.autograd.graph.increment_version(self._table)
        self._table._cednerf_f16.adopt(self._table)

    def _use_nvls(self) -> bool:
        """In-switch reduce / broadcast pays from about six ranks on: at N = 2 every element still crosses the requester's
        links twice (measured 0.55 ms against 0.24 ms for peer loads / stores), at N = 8 it replaces seven inbound copies
        by one (0.34 against 0.47 ms).  `nvls` = True / False forces it; the default "auto" decides by world size."""
        if self._peer is None or not self._peer.mc_base or self.replicate_master:
            return False
        return self.world >= 6 if self.nvls == "auto" else bool(self.nvls)

    def transport(self) -> str:
        """How the table's reduce-scatter / all-gather travels: 'nvls' (in-switch), 'p2p-symm' or 'p2p-ipc'."""
        if self._peer is None:
            return "none"
        return "nvls" if self._use_nvls() else "p2p-" + self._peer.provider

    def timed_out(self) -> bool:
        """Host read: did a barrier ever give up waiting for a peer?"""
        return self._peer is not None and int(self._ctrl_i32[9].item()) != 0

    def close(self):
        if self._peer is not None:
            torch.cuda.synchronize()
            if getattr(self, "_ops", None) is not None:
                self._ops.set_table_grad_alloc(None)
            self._peer.close()


def ctypes_byref(obj):
    import ctypes

    return ctypes.byref(obj)


class SharedRng:
    """`rng` argument of OccGridEstimator.update_every_n_steps for data-parallel training: every rank draws the cells,
    the jitter and (through `rand`) the occ_eval_fn timestamps from a device generator seeded alike, so that the
    occupancy replicas - functions of identical parameters at identical points - stay bit-identical without a
    collective.  (`sync_occupancy` is the belt-and-braces alternative: an all-reduce(MAX) of `occs`.)"""

    def __init__(self, seed: int, device):
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))

    def randint(self, high: int, n: int) -> torch.Tensor:
        return torch.randint(int(high), (int(n),), generator=self.gen, device=self.device)

    def rand(self, *shape) -> torch.Tensor:
        return torch.rand(*shape, generator=self.gen, device=self.device)


def sync_occupancy(estimator, occ_thre: float = 1e-2, group=None) -> None:
    """all-reduce(MAX) of `occs` followed by the re-threshold of nerfacc's `_update` tail: replicas that were updated
    from different draws agree again (cells one rank saw as occupied stay occupied everywhere)."""
    from . import ops
    from .nerfacc.grid import set_occupancy_bits

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(estimator.occs, op=dist.ReduceOp.MAX, group=group)
    valid = estimator.occs >= 0
    thre = torch.clamp((estimator.occs * valid).sum() / valid.sum().clamp_min(1), max=occ_thre).reshape(1).contiguous()
    bits = torch.empty(estimator.occs.numel() // 32, dtype=torch.int32, device=estimator.occs.device)
    ops.occ_threshold_pack(estimator.occs, thre, estimator.binaries, bits)
    set_occupancy_bits(estimator.binaries, bits)
