#!/usr/bin/env python
"""Benchmark of the physics-guided sampler step (BASELINE.json metric: guided sample-steps/s; guidance kernel HBM GB/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU torch sampler (oracle port)

A "step" is ONE guided Heun step (two denoiser evaluations with the finite-difference time derivative, the three
guidance losses, the gradient through the denoiser(s), the fused update) over one batch.  Workload: config 2 of
BASELINE.json -- heat equation 128x128, batch 512 sharded over 8 GPUs = 64 samples per GPU; N GPUs run 64 N
samples (weak scaling, N = 8 is the named configuration).  Denoiser: the reference's unet-v2 architecture with
seeded random weights (no checkpoints ship with the reference); data: synthetic.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for what every key means.
"""
from __future__ import annotations

import argparse
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


def pde_plugins(pde, ours: bool, dx):
    if ours:
        import dynamical_pde_diffusion_b200 as dp
        return (dp.heat_loss2 if pde == "heat" else dp.llg_residual_loss, {"dx": dx}, dp.X_and_dXdt_fd)
    from oracle import guided_sampler_ref as R
    return (R.heat_loss2 if pde == "heat" else R.llg_residual_loss, {"dx": dx}, R.X_and_dXdt_fd)


# algorithmic bytes per element / pixel of each kernel (DESIGN.md section "Kernels"; fp32 fields, fp64 state)
def algorithmic_bytes(name, B, C_, ch_a, H, W, dummy_dxdt):
    n, px, cu = B * C_ * H * W, B * H * W, C_ - ch_a
    return {
        "dpde_sampler_init": 20 * n,                                   # latents 8 -> x64 8 + x32 4
        "dpde_euler_predict": 16 * n,                                  # x_cur 8 + x0 4 -> x_eu32 4
        "dpde_euler_predict_bwd": 8 * n,                               # g 4 -> seed 4
        "dpde_guidance_reduce": 4 * px * (C_ + (0 if dummy_dxdt else cu)),           # x0 (C) + dxdt_u (C_u)
        "dpde_guidance_vjp": 4 * px * (2 * C_ + (0 if dummy_dxdt else cu)),          # + g (C): SURVEY 8d "(2C+C_u)*4"
        "dpde_heun_guided_update": 36 * n,                             # x_cur 8, x0c 4, x0n 4, g_eu 4, g_cur 4 -> 8 + 4
    }.get(name)


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of JointSampler.sample on the host cores
# ---------------------------------------------------------------------------------------------------------
def time_reference_steps(workload, batch, steps, warmup, seed=0):
    """Time `steps` guided steps of the reference algorithm (oracle restatement, torch CPU, all host threads)."""
    from oracle import guided_sampler_ref as R

    pde, C_, ch_a, H, W, _, n_cfg = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count() or 1)
    net, prob = build_problem(workload, batch, seed)
    loss_fn, loss_kwargs, provider = pde_plugins(pde, False, prob["dx"])
    dev = torch.device("cpu")
    sig = R.karras_sigmas(n_cfg, 0.002, 80.0, 7.0, dev, net)
    F64 = torch.float64
    oa, ou = prob["obs_a"].to(F64), prob["obs_u"].to(F64)
    ma, mu = prob["mask_a"].to(F64), prob["mask_u"].to(F64)
    g = torch.Generator().manual_seed(seed + 7)
    x = torch.randn((batch, C_, H, W), generator=g, dtype=F64) * sig[0]
    i = 0
    for _ in range(warmup):
        x, _ = R.guided_step(net, x, i, sig, prob["labels"], oa, ou, ma, mu, ch_a, loss_fn, loss_kwargs, prob["zeta_a"],
                             prob["zeta_u"], prob["zeta_pde"], n_cfg, provider)
        i += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        x, _ = R.guided_step(net, x, i, sig, prob["labels"], oa, ou, ma, mu, ch_a, loss_fn, loss_kwargs, prob["zeta_a"],
                             prob["zeta_u"], prob["zeta_pde"], n_cfg, provider)
        i += 1
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                                     # only rank 0 times the CPU reference
    pde, C_, ch_a, H, W, b_gpu, n_cfg = WORKLOADS[args.workload]
    batch = args.ref_batch
    value, dt, cores = time_reference_steps(args.workload, batch, args.steps, args.warmup)
    sample = (f"{args.steps} guided steps (after {args.warmup} warm-up) of batch {batch} out of the workload's "
              f"{b_gpu} per GPU, same grid/denoiser/plug-ins, torch {torch.__version__} CPU, fp32 denoiser + fp64 guidance")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 guidance / f32 denoiser", "data": "synthetic",
        "config": config_dict(args, batch_per_gpu=b_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
            "parallelism": f"batch-shard x{args.gpus}, independent shards",
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
    if not args.ieee:                                              # what sampling_context does (sample.py:626-630)
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
    else:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    pde, C_, ch_a, H, W, b_gpu, n_cfg = WORKLOADS[args.workload]
    B = args.batch_per_gpu or b_gpu
    net, prob = build_problem(args.workload, B, seed=rank)
    net = net.to(dev)
    if args.channels_last:                                         # PyTorch-level layout choice for the denoiser only
        net = net.to(memory_format=torch.channels_last)
    loss_fn, loss_kwargs, provider = pde_plugins(pde, True, prob["dx"])
    if args.fd_batched:                                            # opt-in: the two offset evaluations as one 2B-batch call
        provider = dp.X_and_dXdt_fd_batched
    smp = dp.JointSampler(net, dev, (H, W), C_, B, ch_a, loss_fn, loss_kwargs, num_steps=n_cfg, out_and_grad_fn=provider)
    z = (prob["zeta_a"], prob["zeta_u"], prob["zeta_pde"])
    host = {k: prob[k].pin_memory() for k in ("labels", "obs_a", "obs_u", "mask_a", "mask_u")}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
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
    dom = max(kernels, key=lambda k: kernels[k]["avg_us"] * kernels[k]["launches"]) if kernels else None
    roofline = None
    if dom:
        k = kernels[dom]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")     # ncu dram bytes per launch of this workload
        if args.workload == "heat128" and B == 64 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": k["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": round(k["achieved_gbs"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "note": "bench-workload launch (L2-resident, launch-latency regime); see roofline_large_grid"}

    # ---- the same kernels on a grid far larger than L2 (config 5 shape on one GPU) -------------------------
    large = large_grid_rooflines(dev, peak) if rank == 0 and not args.skip_large else None

    # ---- end to end through the public API: host tensors in, host tensors out -------------------------------
    e2e = None
    if not args.skip_e2e:
        est = ms_max / args.steps / 1e3
        n_e2e = args.e2e_steps or max(20, min(n_cfg, int(args.e2e_budget / max(est, 1e-6))))
        barrier()
        t0 = time.perf_counter()
        # every rank samples ITS shard (the host tensors above are already per-rank), then the one collective of
        # the path: an NCCL all-gather of samples + loss traces
        x, tr = smp.sample(host["labels"], host["obs_a"], host["obs_u"], host["mask_a"], host["mask_u"], *z,
                           return_losses=True, num_steps=n_e2e)
        if world > 1:
            x, tr = D.gather_samples(x, tr, B * world, device=dev)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = B * C_ * H * W * 4 + n_e2e * 4 * 4
        e2e = {"value": world * B * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / n_e2e, "d2h_bytes_per_step": d2h / n_e2e,
               "sampler_steps": n_e2e, "seconds": round(dt, 3),
               "api": "JointSampler.sample(host tensors) -> (cpu samples, numpy loss trace)" + (" + NCCL all_gather" if world > 1 else "")}
        assert torch.isfinite(x).all()

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample -----------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        v, dt, cores = time_reference_steps(args.workload, args.ref_batch, args.cpu_steps, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} guided steps (1 warm-up) of batch {args.ref_batch} of the same workload on the host CPU, {dt:.1f} s"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 guidance+state / f32 denoiser", "data": "synthetic", "config": config_dict(args, B),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "roofline_large_grid": large, "kernels": kernels, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def large_grid_rooflines(dev, peak, B=8, H=4096, W=4096, reps=5):
    """Kernel-only timing on config 5's shape (8 x 2 x 4096^2: 1 GiB per fp32 field, far beyond L2)."""
    import ctypes as C
    from dynamical_pde_diffusion_b200 import GuidanceEngine, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    C_, ch_a = 2, 1
    n = B * C_ * H * W
    s = torch.cuda.current_stream().cuda_stream
    x0 = torch.randn(B, C_, H, W, device=dev)
    dxdt = torch.randn(B, C_, H, W, device=dev)
    mask = (torch.rand(H, W, device=dev) < 0.2)
    obs = torch.randn(1, 1, H, W, device=dev)
    eng = GuidanceEngine(B, C_, ch_a, H, W, PDE_HEAT, dev, obs_a=obs, mask_a=mask, obs_u=obs, mask_u=mask,
                         sample_coef=torch.rand(B, device=dev).double(), dx=1.0 / (H - 1))
    x64 = torch.randn(B, C_, H, W, device=dev, dtype=torch.float64)
    o64, o32 = torch.empty_like(x64), torch.empty_like(x0)
    g1, g2 = torch.randn_like(x0), torch.randn_like(x0)
    w = (20.0, 0.5, 20.0)

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best.append(a.elapsed_time(b))
        return sum(best) / len(best)

    out = {}
    t = timeit(lambda: eng.reduce(x0, dxdt, w))
    out["dpde_guidance_reduce"] = (4 * B * H * W * (C_ + 1), t)
    g_holder = {}
    def vjp():
        g_holder["g"] = None
        g_holder["g"] = eng.vjp(x0, dxdt, w)
    t = timeit(vjp)
    out["dpde_guidance_vjp"] = (4 * B * H * W * (2 * C_ + 1), t)
    t = timeit(lambda: _ffi.call("dpde_euler_predict", x64.data_ptr(), x0.data_ptr(), 3.0, 2.0, o32.data_ptr(), n, s))
    out["dpde_euler_predict"] = (16 * n, t)
    t = timeit(lambda: _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0.data_ptr(), dxdt.data_ptr(), g1.data_ptr(),
                                 g2.data_ptr(), 3.0, 2.0, o64.data_ptr(), o32.data_ptr(), n, s))
    out["dpde_heun_guided_update"] = (36 * n, t)
    return {"shape": [B, C_, H, W], "peak": peak, "unit": "GB/s", "timing": f"CUDA events, mean of {reps} launches after 1 warm-up, inputs 1-2 GiB each (>> L2)",
            "kernels": {k: {"algorithmic_bytes": nb, "ms": round(ms, 4), "achieved": round(nb / ms / 1e6, 1), "frac": round(nb / ms / 1e6 / peak, 4)}
                        for k, (nb, ms) in out.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="heat128", choices=sorted(WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--ref-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--e2e-budget", type=float, default=45.0, help="seconds the end-to-end sample() call may take")
    ap.add_argument("--ieee", action="store_true", help="IEEE fp32 convolutions instead of the reference's TF32 setting")
    ap.add_argument("--channels-last", action="store_true", help="run the PyTorch denoiser in NHWC memory format")
    ap.add_argument("--fd-batched", action="store_true", help="finite-difference offsets as one denoiser call of batch 2B")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-large", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
