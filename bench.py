#!/usr/bin/env python
"""Benchmark of the per-ray rendering hot path: DyNeRF flame_salmon_1-shaped train step (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement of the reference path (oracle/)

One step = the body of the reference's training loop for this path (train_real.py:330-420 without the data loader):
estimator.update_every_n_steps (works every 16th step; on a twin estimator, see OccupancyUpdate), stratified
occupancy-grid sampling with the no-grad density pre-pass and visibility filtering, the field forward on the surviving
samples, compositing, the MSE + auxiliary losses of the canonical DyNeRF flags (-te -ta -df -f -wr -ae), backward,
GradScaler (2^10) and fused Adam, at the reference schedule's first-iteration learning rate (see TrainState).  The
reference arm runs the same step on the CPU restatement (oracle/) on a bounded ray sample, without the occupancy
update (8.4 M CPU field queries per update would only flatter the ratio).
Rank 0 prints ONE JSON line (see the keys below).  Synthetic data: seeded rays / pixels / occupancy, random-init
weights with a density boost (cednerf_b200/workload.py)."""
from __future__ import annotations

import argparse
import json
import os

# Sample-sized buffers change size a little every step (the visible-sample count follows the field, and the named
# configuration sits right at 2^20 samples - a size-class boundary of any power-of-two rounding).  The library allocates
# them at sticky capacities (ops._sempty), so the caching allocator sees the same request sizes step after step; 1/8
# power-of-two rounding keeps the few remaining torch-side temporaries in one class.  (Measured over 120 steps: no step
# above 6.3 ms except the occupancy updates; expandable segments, used earlier in the round, map GB-sized growth in
# 2 MB granules and cost one 27-280 ms step whenever a capacity had to grow.)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "roundup_power2_divisions:8")
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=2 ** 18, help="rays per rank per step")
    ap.add_argument("--cpu-rays", type=int, default=4096, help="rays per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=3, help="instrumented steps for the per-kernel breakdown")
    ap.add_argument("--render-frames", type=int, default=2, help="full frames per rank for the render leg (0 = skip)")
    ap.add_argument("--torch-profile", default="", help="write a torch.profiler kernel table of one step to this file")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# the step (shared by both arms: `impl` is cednerf_b200 or the oracle namespace below)
# --------------------------------------------------------------------------------------------------------------
def aux_losses(rgb, acc, pixels, extra, flags):
    """train_real.py:369-409 for the canonical DyNeRF flags (-ae, -wr, -f)."""
    loss = 0.0
    t_last = (1 - acc).clamp(1e-6, 1 - 1e-6)
    loss = loss + (-(t_last * torch.log(t_last) + (1 - t_last) * torch.log(1 - t_last)).mean()) * 1e-3
    for ex in extra:
        rgbper = (ex["rgbs"] - pixels[ex["ray_indices"]]).pow(2).sum(dim=-1)
        loss = loss + (rgbper * ex["weights"].detach()).sum() / pixels.shape[0] * 1e-3
        if flags.get("use_feat_predict"):
            loss = loss + ex["latent_losses"].mean()
    return loss


def make_scheduler(opt, max_steps=20000):
    """train_real.py:276-287: linear warm-up from 0.01 x lr over 100 iterations, then step decay."""
    return torch.optim.lr_scheduler.ChainedScheduler([
        torch.optim.lr_scheduler.LinearLR(opt, start_factor=0.01, total_iters=100),
        torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[max_steps // 2, max_steps * 3 // 4, max_steps * 9 // 10],
                                             gamma=0.33)])


class TrainState:
    """Parameters at iteration 0, so that every timed leg measures the same iterations of training on the same field.
    The learning rate is held at the reference schedule's FIRST-iteration value (LinearLR start_factor 0.01 -> 1e-4):
    random-init weights fitted to random pixels move the density by e-folds within tens of Adam steps at the full rate,
    and the visible-sample count (the work per step) would then be whatever the step count made it - 1.0 M samples at
    step 0, 1.4 M fifteen steps later - instead of the named configuration's 2^20 per 2^18 rays.  Every kernel of the
    step, the optimiser included, does the same work at any learning rate."""

    def __init__(self, field, opt):
        self.field, self.opt = field, opt
        self.params = [p.detach().clone() for p in field.parameters()]
        self.sched = None
        self.restore()

    def restore(self):
        with torch.no_grad():
            for p, q in zip(self.field.parameters(), self.params):
                p.copy_(q)
        self.opt.state.clear()
        if hasattr(self.opt, "_step_t"):
            self.opt._step_t = None
        for g in self.opt.param_groups:
            g["lr"] = g.get("initial_lr", g["lr"])
        make_scheduler(self.opt)  # sets lr to 0.01 x initial_lr (iteration 0 of the reference schedule); never stepped


class OccupancyUpdate:
    """estimator.update_every_n_steps(step, occ_eval_fn, occ_thre) of the reference loop (train_real.py:324-336): every
    16th iteration the field is queried at one jittered point per grid cell (all 4 x 128^3 cells during the first 256
    iterations) with random timestamps.  It runs on a TWIN of the estimator: the synthetic occupancy is what defines the
    workload (a random field would otherwise mark everything occupied), so the marcher keeps reading the original."""

    def __init__(self, impl, cfg, est, field, step_size):
        self.twin = impl.OccGridEstimator(list(cfg.roi_aabb), resolution=cfg.occ_res, levels=cfg.occ_levels)
        self.twin = self.twin.to(est.aabbs.device)
        self.twin.binaries, self.twin.occs = est.binaries.clone(), est.occs.clone()
        self.twin.train()
        self.field, self.step_size, self.step, self.updates = field, step_size, 0, 0

    def occ_eval_fn(self, x):
        t = torch.rand(x.shape[0], 1, device=x.device)
        return self.field.query_density(x, t)["density"] * self.step_size

    def __call__(self):
        if self.step % 16 == 0:
            self.updates += 1
        self.twin.update_every_n_steps(step=self.step, occ_eval_fn=self.occ_eval_fn, occ_thre=1e-2)
        self.step += 1


def train_step(impl, field, est, opt, scaler, batch, cfg, rk, reducer=None, sched=None, occ=None):
    if occ is not None:
        occ()
    rays = impl.Rays(batch["origins"], batch["viewdirs"])
    rgb, acc, depth, n_samples, extra = impl.render_image(field, est, rays, render_bkgd=batch["color_bkgd"],
                                                          timestamps=batch["timestamps"], jitter=batch["jitter"], **rk)
    if n_samples == 0:
        return None, 0
    if hasattr(impl, "losses"):  # this repo: the same loss as one fused forward / backward launch
        loss = impl.losses.training_loss(rgb, acc, batch["pixels"], extra, acc_entropy_loss=True, weight_rgbper=True,
                                         use_feat_predict=bool(cfg.flags.get("use_feat_predict")))
    else:
        loss = torch.nn.functional.mse_loss(rgb, batch["pixels"]) + aux_losses(rgb, acc, batch["pixels"], extra, cfg.flags)
    opt.zero_grad()
    if scaler is not None:
        scaler.scale(loss).backward()
        if reducer is not None:
            reducer.wait()
        scaler.step(opt)
        scaler.update()
    else:
        (loss * 1024.0).backward()
        for p in field.parameters():
            if p.grad is not None:
                p.grad.div_(1024.0)
        opt.step()
    if sched is not None:
        sched.step()
    return loss, n_samples


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_mhz, self.run, self.th = index, [], set(), None, False, None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001 - NVML missing: report it, do not fail the benchmark
            return
        names = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown if hasattr(pynvml, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.run = True

        def loop():
            while self.run:
                try:
                    self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.02)

        self.th = threading.Thread(target=loop, daemon=True)
        self.th.start()

    def stop(self):
        self.run = False
        if self.th is not None:
            self.th.join(timeout=1.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# algorithmic bytes per launch of each entry point (SURVEY.md §8d per-unit figures x units of the launch)
def algorithmic_bytes(name, args):
    if name == "cednerf_hashgrid_fwd":
        n, L = args[2], args[4]._obj.n_levels
        return n * (L * 8 * 4 + 12 + L * 4)                     # 512 B table + 12 B xyz + 64 B features (L = 16)
    if name == "cednerf_hashgrid_bwd":
        n, L, is16 = args[2], args[4]._obj.n_levels, args[7]
        b = n * (2 * L * (2 if is16 else 4) + 12)               # dy + xyz
        if args[8]:
            b += n * L * 8 * 8                                  # fp32 table-gradient RMW, counted once (1024 B)
        if args[9]:
            b += n * (L * 8 * 4 + 12)                           # table re-read for dL/dx + dx write
        return b
    if name == "cednerf_field_fwd":
        n, L = args[9], args[14]._obj.levels.n_levels
        return n * (L * 8 * 4 + 16 + 4 + (12 if args[16] else 0))  # 512 B table + packed sample 16 B + sigma (+ rgb)
    if name == "cednerf_mlp_fwd":
        d, n = args[2]._obj, args[3]
        hid = (d.n_layers - 1) * 128 if args[5] else 0
        return n * (d.dim_in[0] * 2 + d.dim_out[d.n_layers - 1] * 2 + hid)
    if name == "cednerf_mlp_bwd":
        d, n = args[4]._obj, args[5]
        return n * (d.dim_in[0] * 2 + (d.n_layers - 1) * 128 + d.dim_out[d.n_layers - 1] * 2 +
                    (d.dim_in[0] * (4 if args[7] else 2) if args[6] else 0))
    if name == "cednerf_march":
        n = args[3]
        return n * 52                                           # 32 B/ray in + 20 B/ray out; per-sample writes added below
    if name == "cednerf_composite_fwd":
        return args[8] * 44 + args[9] * 20
    if name == "cednerf_composite_bwd":
        return args[8] * 60 + args[9] * 20
    if name == "cednerf_adam_step":                             # 16 B read + 12 B written (+ 2 B fp16 copy) per parameter
        t = args[0]._obj
        return sum(t.n[k] * (30 if t.p16[k] else 28) for k in range(t.n_tensors))
    if name == "cednerf_nonfinite_check":
        t = args[0]._obj
        return sum(t.n[k] * 4 for k in range(t.n_tensors))
    return None


def run_ours(args):
    import torch.distributed as dist

    import cednerf_b200 as cb
    from cednerf_b200 import _lib, dp, workload

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload.DYNERF
    rk = workload.render_kwargs(cfg)
    est, field = workload.build_scene(cfg, dev, cb, seed=42)
    est.train(), field.train()
    opt = cb.optim.FusedAdam(field.parameters(), lr=1e-2, eps=1e-15)  # apex.optimizers.FusedAdam, train_real.py:267-270
    scaler = cb.optim.GradScaler(2 ** 10)                              # torch.cuda.amp.GradScaler(2**10), train_real.py:252
    reducer = dp.GradAllReducer(field.parameters(), world) if world > 1 else None
    state = TrainState(field, opt)
    occ = OccupancyUpdate(cb, cfg, est, field, rk["render_step_size"])

    gen = torch.Generator().manual_seed(1000 + rank)  # every rank draws its own slice of the global batch
    n_host = 4
    host = [workload.draw_batch(cfg, args.rays, gen, pin=True) for _ in range(n_host)]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    # allocator warm-up batches: the visible-sample count moves by a few % from step to step and the configuration sits
    # right at 2^20 samples, a size-class boundary of the caching allocator.  Untimed forward+backward passes (no
    # optimiser step: parameters and optimiser state stay untouched) on 1.06 x and 1.12 x the rays, run after the warm-up
    # steps, leave free blocks of the next size classes in the cache, so the timed step in which the count crosses the
    # boundary does not have to grow the pool (one such step took 58-110 ms).  A training run reaches the same state
    # after its first few hundred iterations.
    # (the count grows by ~0.25 % per step at this learning rate: cover the whole run in 6 % increments)
    n_classes = max(2, int((0.003 * (args.steps + max(args.warmup, 3)) + 0.06) / 0.06) + 1)
    oversized = [{k: v.to(dev) for k, v in workload.draw_batch(cfg, int(args.rays * (1.0 + 0.06 * (j + 1))), gen).items()}
                 for j in range(min(n_classes, 12))]

    def touch_size_classes():
        for b in oversized:
            rays_b = cb.Rays(b["origins"], b["viewdirs"])
            rgb, acc, _, n_s, extra = cb.render_image(field, est, rays_b, render_bkgd=b["color_bkgd"],
                                                      timestamps=b["timestamps"], jitter=b["jitter"], **rk)
            loss = torch.nn.functional.mse_loss(rgb, b["pixels"]) + aux_losses(rgb, acc, b["pixels"], extra, cfg.flags)
            opt.zero_grad()
            scaler.scale(loss).backward()
            if reducer is not None:
                reducer.wait()
            opt.zero_grad()

    def step_resident(i):
        return train_step(cb, field, est, opt, scaler, resident[i % n_host], cfg, rk, reducer, state.sched, occ)

    def step_e2e(i):
        b = {k: v.to(dev, non_blocking=True) for k, v in host[i % n_host].items()}
        loss, n_s = train_step(cb, field, est, opt, scaler, b, cfg, rk, reducer, state.sched, occ)
        return (None if loss is None else float(loss.item())), n_s  # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        state.restore()          # iteration 0 again (untimed), then the warm-up steps allocate the optimiser state
        occ.step = 0
        for i in range(max(args.warmup, 3)):
            fn(i)
        touch_size_classes()
        barrier()
        occ.updates = 0
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_tot = 0
        for i in range(k):
            if os.environ.get("BENCH_DEBUG_STEPS"):
                torch.cuda.synchronize()
                t_dbg = time.perf_counter()
                n_i = fn(i)[1]
                torch.cuda.synchronize()
                print(f"[debug] step {i}: {(time.perf_counter() - t_dbg) * 1e3:.2f} ms, {n_i} samples", file=sys.stderr)
                n_tot += n_i
                continue
            n_tot += fn(i)[1]
        e1.record()
        barrier()
        ms = dp.max_over_ranks(e0.elapsed_time(e1) / k, dev)
        return ms, n_tot / k, (_lib.launch_count() - l0) // k

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, samples, launches = timed(step_resident, args.steps)
    occ_updates = occ.updates
    clk = clocks.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    samples_all = dp.sum_over_ranks(samples, dev)

    # ---- per-entry-point breakdown and the roofline of the dominant kernel (CUDA events, same stream) -----------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    roof, breakdown = None, {}
    if args.profile_steps > 0:
        # every rank runs the instrumented steps (they contain the gradient all-reduce); rank 0 keeps the records
        rec = []
        real_call = _lib.call

        def recording_call(name, *a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            real_call(name, *a)
            e.record()
            rec.append((name, a, s, e))

        state.restore()
        step_resident(0)
        for m in (_lib, cb.ops, cb.optim, cb.losses):
            m.call = recording_call
        barrier()
        t0 = time.perf_counter()
        for i in range(args.profile_steps):
            step_resident(i)
        torch.cuda.synchronize()
        prof_ms = (time.perf_counter() - t0) * 1e3 / args.profile_steps
        for m in (_lib, cb.ops, cb.optim, cb.losses):
            m.call = real_call
    if rank == 0 and args.profile_steps > 0:
        agg = {}
        for name, a, s, e in rec:
            t = s.elapsed_time(e)
            by = algorithmic_bytes(name, a)
            d = agg.setdefault(name, {"ms": 0.0, "launches": 0, "bytes": 0, "has_bytes": by is not None})
            d["ms"] += t
            d["launches"] += 1
            d["bytes"] += by or 0
        for name, d in agg.items():
            breakdown[name] = {"ms_per_step": round(d["ms"] / args.profile_steps, 4),
                               "launches_per_step": d["launches"] // args.profile_steps}
        top = max((n for n in agg if agg[n]["has_bytes"]), key=lambda n: agg[n]["ms"])
        d = agg[top]
        achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        traffic = None  # dram__bytes_read + dram__bytes_write per launch of that kernel, from the committed ncu capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(top, {}).get("bytes_per_launch")
        except (OSError, ValueError):
            pass
        roof = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": round(d["ms"] / d["launches"], 4), "bytes_per_launch": d["bytes"] // d["launches"],
                "share_of_step": round(d["ms"] / args.profile_steps / prof_ms, 3)}
        breakdown["_ours_total_ms"] = round(sum(v["ms"] for v in agg.values()) / args.profile_steps, 3)
        breakdown["_instrumented_step_ms"] = round(prof_ms, 3)

    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step_resident(0)
            torch.cuda.synchronize()
        if rank == 0:
            with open(args.torch_profile, "w") as f:
                f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))

    # ---- render leg (BASELINE.json configs[3]): full 1352 x 1014 frames through render_image_test, frames sharded over
    # ranks with no collective; reported beside the train-step headline --------------------------------------------------
    render = None
    if args.render_frames > 0:
        field.eval(), est.eval()
        o, d = workload.frame_rays(cfg, 0)
        frame = cb.Rays(o.to(dev).view(cfg.height, cfg.width, 3), d.to(dev).view(cfg.height, cfg.width, 3))
        black = torch.zeros(3, device=dev)
        frames = dp.shard_interleaved(args.render_frames * world, rank, world)   # weak scaling: render_frames per rank

        def render_one(i):
            t = torch.tensor([[frames[i % len(frames)] / 300.0]], device=dev)     # t = i / 300 (dnerf_3d_video_IS.py:366)
            return cb.render_image_test(1024, field, est, frame, render_bkgd=black, timestamps=t, **rk)[3]

        render_one(0)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        n_render = sum(render_one(i) for i in range(len(frames)))
        r1.record()
        barrier()
        ms_r = dp.max_over_ranks(r0.elapsed_time(r1), dev)
        n_render_all = dp.sum_over_ranks(n_render, dev)
        rays_r = cfg.width * cfg.height * len(frames) * world
        render = {"workload": f"{cfg.width}x{cfg.height} frames, render_image_test(max_samples=1024), "
                              f"{len(frames)} frame(s)/GPU, frame-sharded, no collective",
                  "rays_per_s": round(rays_r / (ms_r * 1e-3), 1), "samples_per_s": round(n_render_all / (ms_r * 1e-3), 1),
                  "ms_per_frame": round(ms_r / len(frames), 3), "samples_per_ray": round(n_render_all / rays_r, 3)}
        if rank == 0 and args.profile_steps > 0:   # per-entry-point share of one frame
            rec_r = []
            real_call = _lib.call

            def rec_call(name, *a2):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record()
                real_call(name, *a2)
                e_.record()
                rec_r.append((name, s_, e_))

            for m in (_lib, cb.ops, cb.optim, cb.losses):
                m.call = rec_call
            render_one(0)
            torch.cuda.synchronize()
            for m in (_lib, cb.ops, cb.optim, cb.losses):
                m.call = real_call
            agg_r = {}
            for name, s_, e_ in rec_r:
                d_ = agg_r.setdefault(name, [0.0, 0])
                d_[0] += s_.elapsed_time(e_)
                d_[1] += 1
            render["breakdown_ms_per_frame"] = {k: [round(v[0], 3), v[1]] for k, v in
                                                sorted(agg_r.items(), key=lambda kv: -kv[1][0])}
        field.train(), est.train()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_reference_steps(args.cpu_rays, steps=3, warmup=1)

    if rank == 0:
        rays_all = args.rays * world
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        line = {
            "metric": "train_step_rays_per_s", "value": round(rays_all / (ms * 1e-3), 1), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 (MLP, tables) / f32 (marching, compositing, grads)",
            "data": "synthetic",
            "config": {"workload": f"{cfg.name} train step, {args.rays} rays/GPU/step, occgrid sampler "
                                   f"(BASELINE.json configs[1]); flags -te -ta -df -f -wr -ae; GradScaler 2^10 + fused Adam (cednerf_b200.optim)",
                       "rays_per_gpu": args.rays, "samples_per_step": round(samples_all, 1),
                       "samples_per_ray": round(samples_all / rays_all, 3),
                       "samples_per_s": round(samples_all / (ms * 1e-3), 1),
                       "l2": "working set (96 MB fp16 table + 191 MB fp32 master + 383 MB Adam state + per-step "
                             "buffers) far exceeds the 126 MB L2; 4 rotating input batches",
                       "occ_update": f"update_every_n_steps(n=16, warm-up: all 4 x 128^3 cells) called every step on a twin "
                                     f"estimator; {occ_updates} update(s) fell into the {args.steps} timed steps",
                       "parallelism": f"dp{world}" if world > 1 else "single"},
            "e2e": {"value": round(rays_all / (ms_e2e * 1e-3), 1), "unit": "rays/s", "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "render": render,
            "breakdown": breakdown,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference path (the reference itself cannot run: nerfacc / tiny-cuda-nn /
# Taichi are CUDA-only third-party packages that are not installed and the reference has no CPU path)
# --------------------------------------------------------------------------------------------------------------
class _OracleImpl:
    def __init__(self):
        from oracle import cednerf_ref as cr
        from oracle import nerfacc_ref as nf

        self.OccGridEstimator, self.DNGPradianceField = nf.OccGridEstimator, cr.DNGPradianceField
        self.Rays, self.render_image = cr.Rays, cr.render_image


def run_reference_steps(n_rays, steps, warmup):
    from cednerf_b200 import workload

    torch.set_num_threads(os.cpu_count() or 1)
    impl = _OracleImpl()
    cfg = workload.DYNERF
    rk = workload.render_kwargs(cfg)
    est, field = workload.build_scene(cfg, "cpu", impl, seed=42)
    est.train(), field.train()
    opt = torch.optim.Adam(field.parameters(), lr=1e-2, eps=1e-15)
    gen = torch.Generator().manual_seed(1000)
    batches = [workload.draw_batch(cfg, n_rays, gen) for _ in range(2)]
    times, samples = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, n_s = train_step(impl, field, est, opt, None, batches[i % 2], cfg, rk)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            samples += n_s
    sec = sum(times) / len(times)
    return {"value": round(n_rays / sec, 1), "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_rays} rays/step of the same workload, {steps} timed steps after {warmup} warm-up "
                      f"({round(samples / steps / n_rays, 2)} samples/ray), oracle/ PyTorch-CPU + OpenMP C marcher",
            "ms_per_step": round(sec * 1e3, 1), "samples_per_s": round(samples / steps / sec, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from cednerf_b200 import workload

    cfg = workload.DYNERF
    cpu = run_reference_steps(args.cpu_rays, steps=max(1, min(args.steps, 3)), warmup=1)
    line = {"impl": "reference", "metric": "train_step_rays_per_s", "value": cpu["value"], "unit": "rays/s",
            "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": cpu["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (CPU), fp16-rounded MLP",
            "data": "synthetic",
            "config": {"workload": f"{cfg.name} train step, bounded sample of {args.cpu_rays} rays/step on the host CPU "
                                   f"(BASELINE.json configs[1]); flags -te -ta -df -f -wr -ae; Adam"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
