"""Micro-benchmark of the fused reduce-scatter + Adam + all-gather kernel (csrc/dp.cu) under torchrun: full step and
its parts (remote gradient reads only / replica stores only / local only) on a 47.9 M-element table."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cednerf_b200 import _lib, dp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = 2 * 23928800
pm = dp.PeerMemory([("grad", 4 * n), ("p32", 4 * n), ("p16", 2 * n)], dev)
if rank == 0:
    print("provider", pm.provider, "multicast", bool(pm.mc_base), flush=True)
q = (n // 4) // world
lo, hi = 4 * q * rank, (n if rank == world - 1 else 4 * q * (rank + 1))
m = torch.zeros(hi - lo, device=dev)
v = torch.zeros(hi - lo, device=dev)
step = torch.ones(1, device=dev)
pm.local("grad", torch.float32).normal_()


def run(read_world, n_out, label, p32_remote=True, nvls=False):
    a = _lib.DpAdam()
    a.world, a.rank = read_world, (rank if read_world == world else 0)
    order = [rank] + [r for r in range(world) if r != rank]
    for k in range(read_world):
        a.grad[k] = pm.address(k if read_world == world else order[k], "grad")
    a.n_out = n_out
    for k in range(n_out):
        a.p32_out[k] = pm.address(order[k], "p32") if (k == 0 or p32_remote) else None
        a.p16_out[k] = pm.address(order[k], "p16")
    a.m, a.v, a.lo, a.hi = m.data_ptr(), v.data_ptr(), lo, hi
    a.lr, a.weight_decay, a.grad_div = 1e-4, 0.0, float(world)
    if nvls:
        a.grad_mc, a.p16_mc = pm.multicast_address("grad"), pm.multicast_address("p16")
    def go():
        _lib.call("cednerf_dp_adam", ctypes.byref(a), step.data_ptr(), None, None, 0.9, 0.999, 1e-15, 1, _lib.stream())
    for _ in range(3):
        go()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nel = hi - lo
    if rank == 0:
        print(f"{label:28s} {ms:7.3f} ms   remote in {4 * nel * (read_world - 1) / ms / 1e6:7.1f} GB/s   "
              f"remote out {(6 if p32_remote else 2) * nel * (n_out - 1) / ms / 1e6:7.1f} GB/s   local {nel * (4 + 20 + 6) / ms / 1e6:7.1f} GB/s", flush=True)
    dist.barrier()


run(1, 1, "local only")
run(world, 1, "remote reads + local")
run(1, world, "replica stores + local")
run(world, world, "full")
run(world, world, "full, fp16 replicas only", False)
if pm.mc_base:
    run(world, world, "NVLS (ld_reduce + multimem.st)", False, True)
# raw P2P bandwidth of this box with a plain SM copy kernel (torch's elementwise copy on UVA pointers)
peer = (rank + 1) % world
nb = 4 * n
remote = torch.as_tensor(dp._RawCuda(pm.address(peer, "grad"), nb, pm), device=dev)
local = torch.empty(nb, dtype=torch.uint8, device=dev)
for label, dst, src in (("P2P read  (peer -> local)", local, remote), ("P2P write (local -> peer)", remote, local),
                        ("local copy", local, pm.local("p32", torch.uint8))):
    for _ in range(2):
        dst.view(torch.float32).copy_(src.view(torch.float32))
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dst.view(torch.float32).copy_(src.view(torch.float32))
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{label:28s} {nb / (e0.elapsed_time(e1) / 10) / 1e6:7.1f} GB/s (both ranks at once)", flush=True)
    dist.barrier()
dist.barrier()
dist.destroy_process_group()
