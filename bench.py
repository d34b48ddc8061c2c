#!/usr/bin/env python
"""Benchmark of the physics-guided sampler step (BASELINE.json metric: guided sample-steps/s; guidance kernel HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the UNMODIFIED reference sampler on the host CPU

A "step" is ONE guided Heun step (two denoiser evaluations with the finite-difference time derivative, the three
guidance losses, the gradient through the denoiser(s), the fused update) over one batch.  Default workload: config 2
of BASELINE.json -- heat equation 128x128, batch 512 sharded over 8 GPUs = 64 samples per GPU; N GPUs run 64 N
samples (weak scaling, N = 8 is the named configuration).  Denoiser: the reference's unet-v2 architecture with
seeded random weights (no checkpoints ship with the reference); data: synthetic.

Besides the contract's keys the line carries (see DESIGN.md section "Measurement"):
  gpu_reference / speedup_vs_gpu_reference   the unmodified reference sampler (oracle/_ref) on the SAME GPU, same batch
  roofline_large_grid                        every hand-written kernel on shapes far larger than L2 (heat 8x2x4096^2, LLG 8x6x2048^2)
  config5_slab                               heat 4096^2, batch 8: whole grid at N = 1, row slabs + NVLink halo at N > 1 (strong scaling)
  coupled                                    (N > 1) batch shards with the per-step cross-rank sum exchange
  config4_sweep                              a bounded slice of the zeta / num-steps sweep dealt over the ranks

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "guided sample-steps/sec"
UNIT = "sample-steps/s"
WORKLOADS = {
    # name: (pde, C, ch_a, H, W, batch per GPU, sampler steps of the named config)
    "heat128": ("heat", 2, 1, 128, 128, 64, 200),      # config 2 (per-GPU shard of batch 512 over 8 GPUs)
    "heat64": ("heat", 2, 1, 64, 64, 4, 20),           # config 1 (the reference's CPU-runnable case)
    "llg128": ("llg", 6, 3, 128, 128, 32, 50),         # config 3 (batch 256 over 8 GPUs)
}
SLAB = dict(B=8, H=4096, W=4096, C=2, ch_a=1, schedule=200)   # config 5


# ---------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every 100 ms DURING the timed region (NVML)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz, self._stop = index, [], set(), None, threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(index).uuid)).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.thread is not None:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def set_conv_precision(ieee: bool):
    """The reference samples inside `sampling_context`, which sets cuDNN convolutions to TF32 (sample.py:626-630);
    the new-style setter is used because mixing it with the legacy allow_tf32 flags raises in torch >= 2.9."""
    torch.backends.cudnn.conv.fp32_precision = "ieee" if ieee else "tf32"


def build_problem(workload, batch, seed):
    from dynamical_pde_diffusion_b200 import synthetic
    from dynamical_pde_diffusion_b200.denoiser import build_unet_v2, randomize_zero_init

    pde, C_, ch_a, H, W, _, _ = WORKLOADS[workload]
    torch.manual_seed(1234)                                   # same weights on every rank and in both arms
    label_dim = 2 if pde == "heat" else 4
    net = build_unet_v2(C_, label_dim).eval()
    randomize_zero_init(net, seed=99)
    prob = synthetic.heat_problem(batch, H, W, seed=seed) if pde == "heat" else synthetic.llg_problem(batch, H, W, seed=seed)
    return net, prob


def reference_sampler(workload, state_dict, prob, device, batch):
    """The UNMODIFIED reference `JointSampler` (oracle/_ref or /root/reference) around the reference's own
    `EDMWrapper(EDMUNet)` carrying the same weights as our arm's denoiser.  Returns (sampler, kind)."""
    from oracle.ref_import import import_reference, reference_kind

    S, PL, M = import_reference()
    pde, C_, ch_a, H, W, _, n_cfg = WORKLOADS[workload]
    label_dim = 2 if pde == "heat" else 4
    net = M.EDMWrapper(unet=M.EDMUNet(img_channels=C_, label_dim=label_dim, obs_channels=0, base_channels=64,
                                      channel_mults=[1, 2, 2], num_res_blocks=2, dropout=0.0, sigma_emb_dim=64, emb_dim=256),
                       sigma_data=0.5).eval()
    net.load_state_dict(state_dict, strict=True)
    if pde == "heat":
        loss_fn, kw = PL.heat_loss2, {"dx": prob["dx"]}
    else:   # the reference never wires its m x H_eff residual into the sampler: the oracle's torch restatement stands in
        from oracle import guided_sampler_ref as R
        loss_fn, kw = R.llg_residual_loss, {"dx": prob["dx"]}
    smp = S.JointSampler(net, device, (H, W), C_, batch, ch_a, loss_fn, kw, num_steps=n_cfg, out_and_grad_fn=S.X_and_dXdt_fd)
    return smp, S, reference_kind()


def pde_plugins(pde, dx):
    import dynamical_pde_diffusion_b200 as dp
    return (dp.heat_loss2 if pde == "heat" else dp.llg_residual_loss, {"dx": dx}, dp.X_and_dXdt_fd)


# algorithmic bytes per element / pixel of each kernel (DESIGN.md section "Kernels"; fp32 fields, fp64 state)
def algorithmic_bytes(name, B, C_, ch_a, H, W, dummy_dxdt):
    n, px, cu = B * C_ * H * W, B * H * W, C_ - ch_a
    return {
        "dpde_sampler_init": 20 * n,                                   # latents 8 -> x64 8 + x32 4
        "dpde_euler_predict": 16 * n,                                  # x_cur 8 + x0 4 -> x_eu32 4
        "dpde_euler_predict_bwd": 8 * n,                               # g 4 -> seed 4
        "dpde_guidance_reduce": 4 * px * (C_ + (0 if dummy_dxdt else cu)),           # x0 (C) + dxdt_u (C_u)
        "dpde_guidance_vjp": 4 * px * (2 * C_ + (0 if dummy_dxdt else cu)),          # + g (C): SURVEY 8d "(2C+C_u)*4"
        "dpde_guidance_seed": 4 * px * (2 * C_ + (0 if dummy_dxdt else cu)),         # single-launch reduce + VJP: fields read once from HBM
        "dpde_heun_guided_update": 36 * n,                             # x_cur 8, x0c 4, x0n 4, g_eu 4, g_cur 4 -> 8 + 4
    }.get(name)


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference JointSampler.sample on the host cores
# ---------------------------------------------------------------------------------------------------------
def time_reference_cpu(workload, batch, steps, warmup, seed=0):
    """One `sample()` call of `steps` guided steps on the CPU (all host threads) after a `warmup`-step call.

    The reference exposes whole `sample()` calls only (sample.py:278-363), so a bounded sample of the workload is a
    call with a short schedule: `steps - 1` Heun steps + the final Euler step, host tensors in and out."""
    from oracle.ref_import import reference_available

    torch.set_num_threads(os.cpu_count() or 1)
    net, prob = build_problem(workload, batch, seed)
    dev = torch.device("cpu")
    z = (prob["zeta_a"], prob["zeta_u"], prob["zeta_pde"])
    args = (prob["labels"], prob["obs_a"], prob["obs_u"], prob["mask_a"], prob["mask_u"], *z)
    if reference_available():
        smp, _, kind = reference_sampler(workload, net.state_dict(), prob, dev, batch)
        call = lambda n: smp.sample(*args, return_losses=True, num_steps=n)
        kind = "reference"
    else:   # no oracle/_ref in this checkout: the torch restatement of the same loop
        from oracle import guided_sampler_ref as R
        pde, C_, ch_a, H, W, _, _ = WORKLOADS[workload]
        loss_fn = R.heat_loss2 if pde == "heat" else R.llg_residual_loss
        call = lambda n: R.joint_sample(net, dev, (H, W), C_, ch_a, loss_fn, {"dx": prob["dx"]}, *args, num_steps=n,
                                        out_and_grad_fn=R.X_and_dXdt_fd)
        kind = "port"
    if warmup > 0:
        call(max(2, warmup))
    torch.manual_seed(seed + 7)
    t0 = time.perf_counter()
    x, tr = call(max(2, steps))
    dt = time.perf_counter() - t0
    assert torch.isfinite(x).all()
    return batch * max(2, steps) / dt, dt, torch.get_num_threads(), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                                     # only rank 0 times the CPU reference
    pde, C_, ch_a, H, W, b_gpu, n_cfg = WORKLOADS[args.workload]
    batch = args.ref_batch
    steps = max(2, args.steps)
    value, dt, cores, kind = time_reference_cpu(args.workload, batch, steps, min(args.warmup, 3))
    sample = (f"one JointSampler.sample() call of {steps} guided steps (after a {max(2, min(args.warmup, 3))}-step warm-up call) on a batch of "
              f"{batch} of the workload's {b_gpu} samples per GPU, same grid / denoiser weights / plug-ins, torch {torch.__version__} CPU "
              f"with {cores} threads, fp32 denoiser + fp64 guidance; the unmodified reference class from oracle/_ref"
              if kind == "reference" else
              f"{steps} guided steps of batch {batch} with the oracle port (oracle/_ref not installed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 guidance / f32 denoiser", "data": "synthetic",
        "config": config_dict(args, batch_per_gpu=b_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def config_dict(args, batch_per_gpu):
    pde, C_, ch_a, H, W, _, n_cfg = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: {pde} {H}x{W}, C={C_}, {n_cfg}-step schedule, batch {batch_per_gpu} per GPU "
                        f"(config 2 = batch 512 over 8 GPUs)" if args.workload == "heat128" else
                        f"{args.workload}: {pde} {H}x{W}, C={C_}, {n_cfg}-step schedule, batch {batch_per_gpu} per GPU",
            "global_batch": batch_per_gpu * args.gpus, "grid": [H, W], "channels": C_, "schedule_steps": n_cfg,
            "denoiser": "unet-v2 (7.0 M params, random init), time derivative by central FD (3 evaluations)",
            "pde_loss": "heat_loss2 (pde_losses.py:71-96)" if pde == "heat" else
                        "LLG m x H_eff residual: exchange + uniaxial anisotropy (K0 = 0 as the reference) + applied field",
            "parallelism": f"batch-shard x{args.gpus}, " + ("coupled shards (per-step sum exchange)" if getattr(args, "coupled", False) else "independent shards"),
            "denoiser_conv_precision": "ieee fp32" if args.ieee else "tf32 (as the reference's sampling_context, sample.py:626-630)",
            "cache": "inputs of every step are freshly produced tensors; per-step working set (denoiser activations, ~1 GiB per sample: 64 GiB at batch 64) exceeds the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200 import _ffi, distributed as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the CUDA kernels are the product, there is no CPU path")
    rank, world, local = D.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    set_conv_precision(args.ieee)
    torch.backends.cudnn.benchmark = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pde, C_, ch_a, H, W, b_gpu, n_cfg = WORKLOADS[args.workload]
    B = args.batch_per_gpu or b_gpu
    net, prob = build_problem(args.workload, B, seed=rank)
    net_cpu_state = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(dev)
    if args.channels_last:                                         # PyTorch-level layout choice for the denoiser only
        net = net.to(memory_format=torch.channels_last)
    loss_fn, loss_kwargs, provider = pde_plugins(pde, prob["dx"])
    if args.fd_batched:                                            # opt-in: the two offset evaluations as one 2B-batch call
        provider = dp.X_and_dXdt_fd_batched
    smp = dp.JointSampler(net, dev, (H, W), C_, B, ch_a, loss_fn, loss_kwargs, num_steps=n_cfg, out_and_grad_fn=provider,
                          coupled=args.coupled and world > 1)
    z = (prob["zeta_a"], prob["zeta_u"], prob["zeta_pde"])
    host = {k: prob[k].pin_memory() for k in ("labels", "obs_a", "obs_u", "mask_a", "mask_u")}

    # ---- device-resident steps: W warm-up, K timed ------------------------------------------------------------
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    smp.begin(host["labels"], host["obs_a"], host["obs_u"], host["mask_a"], host["mask_u"], *z, generator=gen)
    for _ in range(args.warmup):
        smp.step()
    kernel_events = []
    _ffi.event_log = kernel_events                                  # CUDA events around every C-ABI launch
    barrier()
    launches0 = _ffi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        torch.cuda.nvtx.range_push("timed")                       # lets `ncu --nvtx --nvtx-include timed/` see exactly this region
        e0.record()
        for _ in range(args.steps):
            smp.step()
        e1.record()
        barrier()
        torch.cuda.nvtx.range_pop()
    _ffi.event_log = None
    ms = e0.elapsed_time(e1)
    launches = _ffi.launch_count - launches0
    smp.finish()
    ms_max = max_over_ranks(ms)
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- per-kernel durations from the events recorded inside the timed region ------------------------------
    per_kernel = {}
    for name, a, b in kernel_events:
        per_kernel.setdefault(name, []).append(a.elapsed_time(b) * 1e3)      # microseconds
    peak, peak_src = measured_peaks()
    dummy = False                                                  # both workloads use the finite-difference provider
    kernels = {}
    for name, us in per_kernel.items():
        nbytes = algorithmic_bytes(name, B, C_, ch_a, H, W, dummy)
        avg = sum(us) / len(us)
        kernels[name] = {"launches": len(us), "avg_us": round(avg, 2), "share_of_step": round(sum(us) / (ms * 1e3), 5),
                         "algorithmic_bytes": nbytes, "achieved_gbs": round(nbytes / avg / 1e3, 1) if nbytes else None}
    timed = {k: v for k, v in kernels.items() if v["algorithmic_bytes"]}
    dom = max(timed, key=lambda k: timed[k]["avg_us"] * timed[k]["launches"]) if timed else None
    roofline = None
    if dom:
        k = kernels[dom]
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")     # ncu dram bytes per launch of this workload
        if args.workload == "heat128" and B == 64 and os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("kernels", {}).get(dom), tj.get("captured")
        roofline = {"kernel": dom, "bound": "hbm", "achieved": k["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(k["achieved_gbs"] / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "note": "bench-workload launch (8 MB fields, L2-resident, launch-latency regime: 12.6 MB is 1.9 us at peak); "
                            "the HBM-bound measurement of the same kernels is roofline_large_grid"}

    # ---- end to end through the public API: host tensors in, host tensors out -------------------------------
    e2e, n_e2e = None, 0
    if not args.skip_e2e:
        est = ms_max / args.steps / 1e3
        n_e2e = args.e2e_steps or max(20, min(n_cfg, int(args.e2e_budget / max(est, 1e-6))))
        # warm-up call, as the gpu_reference leg does for the reference (allocator blocks, cuDNN plans of the sample() path)
        smp.sample(host["labels"], host["obs_a"], host["obs_u"], host["mask_a"], host["mask_u"], *z, return_losses=True, num_steps=3)
        barrier()
        t0 = time.perf_counter()
        # every rank samples ITS shard (the host tensors above are already per-rank), then the one collective of
        # the path: an NCCL all-gather of samples + loss traces
        x, tr = smp.sample(host["labels"], host["obs_a"], host["obs_u"], host["mask_a"], host["mask_u"], *z,
                           return_losses=True, num_steps=n_e2e)
        if world > 1:
            x, tr = D.gather_samples(x, tr, B * world, device=dev)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = B * C_ * H * W * 4 + n_e2e * 4 * 4
        e2e = {"value": world * B * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / n_e2e, "d2h_bytes_per_step": d2h / n_e2e,
               "sampler_steps": n_e2e, "seconds": round(dt, 3),
               "api": "JointSampler.sample(host tensors) -> (cpu samples, numpy loss trace)" + (" + NCCL all_gather" if world > 1 else "")}
        assert torch.isfinite(x).all()
        del x, tr

    # ---- (N > 1) the same steps with coupled shards: one cross-rank exchange of the three sums per step ---------
    coupled = None
    if world > 1 and not args.coupled and not args.skip_extras:
        cs = dp.JointSampler(net, dev, (H, W), C_, B, ch_a, loss_fn, loss_kwargs, num_steps=n_cfg, out_and_grad_fn=provider, coupled=True)
        cs.begin(host["labels"], host["obs_a"], host["obs_u"], host["mask_a"], host["mask_u"], *z, generator=gen)
        for _ in range(3):
            cs.step()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            cs.step()
        c1.record()
        barrier()
        cms = max_over_ranks(c0.elapsed_time(c1)) / 5
        cs.finish()
        coupled = {"ms_per_step": round(cms, 3), "value": world * B / (cms / 1e3), "unit": UNIT, "steps": 5,
                   "collective": "all-reduce of 3 fp64 partial sums (24 B) between the reduce and the VJP pass of every step",
                   "independent_ms_per_step": round(ms_max / args.steps, 3)}
        del cs

    del smp
    gc.collect()
    torch.cuda.empty_cache()

    # ---- the unmodified reference sampler on the SAME GPU (rank 0, N = 1): the honest baseline of e2e ---------
    gpu_ref = None
    if rank == 0 and world == 1 and not args.skip_gpu_ref and e2e is not None:
        gpu_ref = time_gpu_reference(args, net_cpu_state, prob, dev, B, n_e2e, host, z)

    # ---- the same kernels on grids far larger than L2 (config 5 shape; LLG 8 x 6 x 2048^2) -------------------
    # (after the end-to-end legs: its multi-GiB operands and empty_cache() calls hand the denoiser's activation blocks back to the
    #  driver; the sample() calls timed above should see the allocator state the timed steps left)
    gc.collect()
    torch.cuda.empty_cache()
    large = large_grid_rooflines(dev, peak) if rank == 0 and not args.skip_large else None

    net = net.cpu()
    del net
    gc.collect()
    torch.cuda.empty_cache()

    # ---- config 5: heat 4096^2, batch 8 -- whole grid at N = 1, row slabs + peer-memory halo exchange at N > 1 --
    slab = None
    if not args.skip_extras and not args.skip_slab:
        slab = slab_leg(rank, world, dev, barrier, max_over_ranks)
        gc.collect()
        torch.cuda.empty_cache()

    # ---- config 4: a bounded slice of the zeta / num-steps sweep, work items dealt over the ranks --------------
    sweep = None
    if not args.skip_extras and not args.skip_sweep:
        sweep = sweep_leg(rank, world, dev, barrier, max_over_ranks)
        gc.collect()
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1 only): the unmodified reference on the host cores, bounded sample ----------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, dt, cores, kind = time_reference_cpu(args.workload, args.cpu_batch, args.cpu_steps, 0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"one JointSampler.sample() call of {max(2, args.cpu_steps)} guided steps on a batch of {args.cpu_batch} of the same workload "
                         f"on the host CPU ({cores} threads), {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 guidance+state / f32 denoiser", "data": "synthetic", "config": config_dict(args, B),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu, "gpu_reference": gpu_ref,
                "speedup_vs_gpu_reference": (round(e2e["value"] / gpu_ref["value"], 4) if gpu_ref and e2e and gpu_ref.get("value") else None),
                "roofline_large_grid": large, "kernels": kernels, "config5_slab": slab, "coupled": coupled, "config4_sweep": sweep}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def time_gpu_reference(args, state_dict, prob, dev, B, n_steps, host, z):
    """`JointSampler.sample` of the UNMODIFIED reference on `dev`, inside its own `sampling_context` (TF32 convolutions,
    sample.py:622-637): same weights, same host inputs, same batch and the same number of steps as our `e2e` call."""
    from oracle.ref_import import reference_available

    if not reference_available():
        return {"unavailable": "oracle/_ref is not installed (python oracle/build_ref.py)"}
    smp, S, kind = reference_sampler(args.workload, state_dict, prob, dev, B)
    call = lambda n: smp.sample(host["labels"], host["obs_a"], host["obs_u"], host["mask_a"], host["mask_u"], *z, return_losses=True, num_steps=n)
    try:
        with S.sampling_context(smp):
            if args.ieee:
                torch.backends.cudnn.conv.fp32_precision = "ieee"
            call(3)                                                 # warm-up: cuDNN autotuning, allocator
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x, tr = call(n_steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        assert torch.isfinite(x).all() and np.isfinite(tr).all()
        return {"value": B * n_steps / dt, "unit": UNIT, "sampler_steps": n_steps, "seconds": round(dt, 3), "batch": B, "kind": kind,
                "ms_per_step": round(1e3 * dt / n_steps, 3),
                "api": "diffusion_pde.sampling.JointSampler.sample (unmodified, device='cuda') inside sampling_context, "
                       "heat_loss2 + X_and_dXdt_fd of the reference, host tensors in / out"}
    except torch.OutOfMemoryError as e:                             # report, never hide
        return {"unavailable": f"reference sampler ran out of device memory at batch {B}: {str(e)[:120]}"}
    finally:
        set_conv_precision(args.ieee)


def _events_ms(fn, reps):
    """Mean CUDA-event duration of `fn` over `reps` launches after one warm-up launch."""
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def large_grid_rooflines(dev, peak, reps=5):
    """Kernel-only timing on shapes far beyond L2: heat on config 5's 8 x 2 x 4096^2 (1 GiB per fp32 field), the LLG
    residual / soft-norm kernels on 8 x 6 x 2048^2 (0.8 GiB per field)."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT, PDE_LLG_NORM, PDE_LLG_RESIDUAL

    out = {}
    s = torch.cuda.current_stream().cuda_stream
    w = (20.0, 0.5, 20.0)

    def guidance(tag, kind, B, C_, ch_a, H, W, coef, dx, with_d):
        cu = C_ - ch_a
        x0 = torch.randn(B, C_, H, W, device=dev)
        dxdt = torch.randn(B, C_, H, W, device=dev) if with_d else None
        mask = (torch.rand(H, W, device=dev) < 0.2)
        obs_a, obs_u = torch.randn(1, ch_a, H, W, device=dev), torch.randn(1, cu, H, W, device=dev)
        eng = GuidanceEngine(B, C_, ch_a, H, W, kind, dev, obs_a=obs_a, mask_a=mask, obs_u=obs_u, mask_u=mask, sample_coef=coef, dx=dx,
                             llg=LLGConstants() if kind == PDE_LLG_RESIDUAL else None)
        px, nd = B * H * W, (cu if with_d else 0)
        out[f"dpde_guidance_reduce[{tag}]"] = (4 * px * (C_ + nd), _events_ms(lambda: eng.reduce(x0, dxdt, w), reps), [B, C_, H, W])
        hold = {}

        def vjp():
            hold["g"] = None
            hold["g"] = eng.vjp(x0, dxdt, w)
        out[f"dpde_guidance_vjp[{tag}]"] = (4 * px * (2 * C_ + nd), _events_ms(vjp, reps), [B, C_, H, W])
        return x0, dxdt

    B, H, W, C_ = 8, 4096, 4096, 2
    x0, dxdt = guidance("heat", PDE_HEAT, B, C_, 1, H, W, torch.rand(B, device=dev).double(), 1.0 / (H - 1), True)
    n = B * C_ * H * W
    x64 = torch.randn(B, C_, H, W, device=dev, dtype=torch.float64)
    o64, o32 = torch.empty_like(x64), torch.empty_like(x0)
    g1, g2 = torch.randn_like(x0), torch.randn_like(x0)
    shp = [B, C_, H, W]
    out["dpde_euler_predict"] = (16 * n, _events_ms(lambda: _ffi.call("dpde_euler_predict", x64.data_ptr(), x0.data_ptr(), 3.0, 2.0, o32.data_ptr(), n, s), reps), shp)
    out["dpde_euler_predict_bwd"] = (8 * n, _events_ms(lambda: _ffi.call("dpde_euler_predict_bwd", g1.data_ptr(), 3.0, 2.0, o32.data_ptr(), n, s), reps), shp)
    out["dpde_heun_guided_update"] = (36 * n, _events_ms(lambda: _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0.data_ptr(), dxdt.data_ptr(), g1.data_ptr(),
                                                                           g2.data_ptr(), 3.0, 2.0, o64.data_ptr(), o32.data_ptr(), n, s), reps), shp)
    out["dpde_sampler_init"] = (20 * n, _events_ms(lambda: _ffi.call("dpde_sampler_init", x64.data_ptr(), 80.0, o64.data_ptr(), o32.data_ptr(), n, s), reps), shp)
    del x0, dxdt, x64, o64, o32, g1, g2
    torch.cuda.empty_cache()
    B, H, W, C_ = 8, 2048, 2048, 6
    guidance("llg_residual", PDE_LLG_RESIDUAL, B, C_, 3, H, W, (1e4 * torch.randn(B, 3, device=dev)).double(), 500e-9 / 64, True)
    guidance("llg_norm", PDE_LLG_NORM, B, C_, 3, H, W, None, 0.0, False)
    torch.cuda.empty_cache()
    return {"peak": peak, "unit": "GB/s", "timing": f"CUDA events, mean of {reps} launches after 1 warm-up, operands 0.8-2 GiB each (>> 126 MB L2)",
            "kernels": {k: {"shape": shp, "algorithmic_bytes": nb, "ms": round(ms, 4), "achieved": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / peak, 4)}
                        for k, (nb, ms, shp) in out.items()}}


def slab_leg(rank, world, dev, barrier, max_over_ranks, steps=6, warmup=2, check_steps=3):
    """Config 5 (heat 4096 x 4096, batch 8; strong scaling): guided steps of a 200-step schedule with the pointwise
    stand-in denoiser (the U-Net cannot run at this size).  N = 1 runs the whole grid on one GPU;
    N > 1 cuts it into row slabs with the peer-memory halo exchange.  Also re-runs a short schedule on rank 0 on the
    whole grid and compares it with the gathered slab result."""
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200 import _ffi, synthetic
    from dynamical_pde_diffusion_b200.denoiser import PointwiseDenoiser
    from dynamical_pde_diffusion_b200.slab import SlabJointSampler, SlabPlan

    B, H, W, C_, ch_a = SLAB["B"], SLAB["H"], SLAB["W"], SLAB["C"], SLAB["ch_a"]
    prob = synthetic.heat_problem(B, H, W, seed=11)                  # same global problem on every rank
    net = PointwiseDenoiser().to(dev)
    z = (prob["zeta_a"], prob["zeta_u"], prob["zeta_pde"])
    args = (prob["labels"], prob["obs_a"], prob["obs_u"], prob["mask_a"], prob["mask_u"], *z)
    kw = dict(num_steps=SLAB["schedule"], out_and_grad_fn=dp.X_and_dXdt_fd)
    if world == 1:
        smp = dp.JointSampler(net, dev, (H, W), C_, B, ch_a, dp.heat_loss2, {"dx": prob["dx"]}, **kw)
    else:
        smp = SlabJointSampler(net, dev, (H, W), C_, B, ch_a, dp.heat_loss2, {"dx": prob["dx"]}, plan=SlabPlan(H, world, rank),
                               transport="peer", **kw)
    gen = torch.Generator(device=dev).manual_seed(5)
    smp.begin(*args, generator=gen)
    for _ in range(warmup):
        smp.step()
    barrier()
    l0 = _ffi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        smp.step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    launches = (_ffi.launch_count - l0) // steps
    # where the step goes: CUDA events around each of OUR launches over two more steps; the rest is the PyTorch stand-in
    # denoiser (6 pointwise evaluations + 2 backwards, ~10 elementwise ATen kernels each) and launch gaps
    ev = []
    _ffi.event_log = ev
    for _ in range(2):
        smp.step()
    torch.cuda.synchronize()
    _ffi.event_log = None
    per = {}
    for name, a, b in ev:
        per[name] = per.get(name, 0.0) + a.elapsed_time(b) / 2
    ours_ms = sum(per.values())
    smp.finish()
    out = {"workload": f"heat {H}x{W}, batch {B}, C={C_}, pointwise stand-in denoiser, FD time derivative, steps of a {SLAB['schedule']}-step schedule",
           "n_gpus": world, "scaling": "strong", "decomposition": "whole grid" if world == 1 else f"{world} row slabs of {H // world} rows + 2 ghost rows per side",
           "ms_per_step": round(ms, 3), "value": B / (ms / 1e3), "unit": UNIT, "pixel_steps_per_s": B * H * W / (ms / 1e3),
           "steps": steps, "warmup": warmup, "our_launches_per_step": launches,
           "our_kernels_ms_per_step": {k: round(v, 3) for k, v in per.items()}, "our_kernels_share": round(ours_ms / ms, 3),
           "limiter": "the PyTorch pointwise stand-in denoiser and its autograd (elementwise ATen kernels over 1-2 GiB tensors) take the "
                      "remaining share of the step; of our launches the fused update + halo push dominates (HBM streaming at ~0.9 of peak)"}
    if world > 1:
        # parity of the decomposition: a short schedule, slabs gathered on every rank vs the whole grid on rank 0
        x, tr = smp.sample(*args, return_losses=True, num_steps=check_steps, generator=torch.Generator(device=dev).manual_seed(9), gather=True)
        if rank == 0:
            whole = dp.JointSampler(net, dev, (H, W), C_, B, ch_a, dp.heat_loss2, {"dx": prob["dx"]}, **kw)
            xr, trr = whole.sample(*args, return_losses=True, num_steps=check_steps, generator=torch.Generator(device=dev).manual_seed(9))
            out["vs_whole_grid"] = {"steps": check_steps, "bit_identical": bool(torch.equal(x, xr) and np.array_equal(tr, trr)),
                                    "max_rel_diff_samples": float((x - xr).abs().max() / xr.abs().max()),
                                    "max_rel_diff_trace": float(np.abs(tr - trr).max() / np.abs(trr).max())}
            del whole, xr
        del x
        barrier()
        smp.release()                                                # unmap the neighbours, then free the exportable buffers
        barrier()
    return out


def sweep_leg(rank, world, dev, barrier, max_over_ranks):
    """Config 4 (zeta / num-steps sensitivity sweep, heat 64 x 64): a bounded slice -- 8 zeta triples x {20} steps x 64
    samples in chunks of 32 = 16 work items -- dealt round-robin over the ranks by `distributed.run_sweep`; fixed total
    work, so the N-GPU value is a strong-scaling point.  The full sweep (4096 samples x {20, 50, 200} x 8) is the same
    call with larger arguments."""
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200 import distributed as D, synthetic
    from dynamical_pde_diffusion_b200.denoiser import build_unet_v2, randomize_zero_init

    H = W = 64
    torch.manual_seed(1234)
    net = build_unet_v2(2, 2).eval()
    randomize_zero_init(net, seed=99)
    net = net.to(dev)
    prob = synthetic.heat_problem(1, H, W, seed=3)
    # 8 zeta_a values log-spaced over the range explored by notebooks/sampler_hyperparameter_opt.ipynb (cell 16)
    zetas = [(float(za), 0.5, 20.0) for za in np.logspace(0, np.log10(2e4), 8)]
    steps, total, chunk = (20,), 64, 32
    make = lambda n, N: dp.JointSampler(net, dev, (H, W), 2, n, 1, dp.heat_loss2, {"dx": prob["dx"]}, num_steps=N)
    problem = {k: prob[k] for k in ("labels", "obs_a", "obs_u", "mask_a", "mask_u")}
    D.run_sweep(make, problem, zetas[:1], (3,), chunk, chunk)        # warm-up (cuDNN autotune for this shape), untimed
    barrier()
    t0 = time.perf_counter()
    final, n_done = D.run_sweep(make, problem, zetas, steps, total, chunk, seed=1)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    total_steps = len(zetas) * sum(steps) * total
    assert np.isfinite(final).all()
    return {"workload": f"heat {H}x{W}, unet-v2, {len(zetas)} zeta_a values x steps {list(steps)} x {total} samples (chunks of {chunk}): "
                        f"{len(zetas) * len(steps) * (total // chunk)} sample() calls dealt over {world} rank(s)",
            "n_gpus": world, "scaling": "strong", "seconds": round(dt, 3), "value": total_steps / dt, "unit": UNIT,
            "sample_steps": total_steps, "collective": "one all-reduce of the (8 x 1 x 5) summary at the end",
            "loss_comb_by_zeta_a": [round(float(v), 4) for v in final[:, 0, 3]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heat128", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--coupled", action="store_true", help="batch shards coupled by the per-step exchange of the three sums (N > 1)")
    ap.add_argument("--ref-batch", type=int, default=16, help="batch of the CPU reference arm's bounded sample")
    ap.add_argument("--cpu-batch", type=int, default=8, help="batch of the cpu_baseline leg inside our arm")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--e2e-budget", type=float, default=25.0, help="seconds the end-to-end sample() call may take")
    ap.add_argument("--ieee", action="store_true", help="IEEE fp32 convolutions instead of the reference's TF32 setting")
    ap.add_argument("--channels-last", action="store_true", help="run the PyTorch denoiser in NHWC memory format")
    ap.add_argument("--fd-batched", action="store_true", help="finite-difference offsets as one denoiser call of batch 2B")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true")
    ap.add_argument("--skip-gpu-ref", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="no config5_slab / coupled / config4_sweep legs")
    ap.add_argument("--skip-slab", action="store_true")
    ap.add_argument("--skip-sweep", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
