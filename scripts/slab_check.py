"""Multi-process row-slab run (one rank per GPU under torchrun): parity against the whole-grid sampler on rank 0 and,
with --bench, the throughput of config 5's shape.  Usage:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29631 \
        scripts/slab_check.py [--transport peer|dist] [--bench --H 4096 --W 4096 --B 8 --steps 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynamical_pde_diffusion_b200 as dp  # noqa: E402
from dynamical_pde_diffusion_b200 import distributed as D  # noqa: E402
from dynamical_pde_diffusion_b200.slab import PointwiseDenoiser, SlabJointSampler, SlabPlan  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", default="peer", choices=["peer", "dist"])
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--H", type=int, default=96)
    ap.add_argument("--W", type=int, default=64)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B, H, W = args.B, args.H, args.W
    g = torch.Generator().manual_seed(0)
    labels = torch.stack([0.5 * torch.rand(B, generator=g), torch.exp(-2.5 + 3 * torch.rand(B, generator=g))], 1).float()
    obs_a, obs_u = torch.randn(1, 1, H, W, generator=g), torch.randn(1, 1, H, W, generator=g)
    mask_a, mask_u = torch.rand(H, W, generator=g) < 0.3, torch.rand(H, W, generator=g) < 0.1
    dx, z = 1.0 / (H - 1), (20.0, 0.5, 20.0)
    net = PointwiseDenoiser().to(dev)
    plan = SlabPlan(H, world, rank)
    smp = SlabJointSampler(net, dev, (H, W), 2, B, 1, dp.heat_loss2, {"dx": dx}, num_steps=max(args.steps, 2),
                           out_and_grad_fn=dp.X_and_dXdt_fd, plan=plan, transport=args.transport)

    if not args.bench:
        lat = torch.randn(B, 2, H, W, generator=g, dtype=torch.float64)
        x, tr = smp.sample(labels, obs_a, obs_u, mask_a, mask_u, *z, return_losses=True, latents=lat, gather=True)
        if rank == 0:
            whole = dp.JointSampler(net, dev, (H, W), 2, B, 1, dp.heat_loss2, {"dx": dx}, num_steps=max(args.steps, 2),
                                    out_and_grad_fn=dp.X_and_dXdt_fd)
            xr, trr = whole.sample(labels, obs_a, obs_u, mask_a, mask_u, *z, return_losses=True, latents=lat)
            ex = float((x - xr).abs().max() / xr.abs().max())
            el = float(np.abs(tr - trr).max() / np.abs(trr).max())
            assert x.shape == xr.shape and ex < 1e-6 and el < 1e-6, (ex, el)
            print(f"slab_check ok: world {world}, transport {args.transport}, sample err {ex:.2e}, trace err {el:.2e}")
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- throughput on a large grid: steps of a 200-step schedule, device-resident, CUDA events, max over ranks ----
    gen = torch.Generator(device=dev).manual_seed(5)
    smp.num_steps = 200
    smp.begin(labels, obs_a, obs_u, mask_a, mask_u, *z, generator=gen)
    for _ in range(args.warmup):
        smp.step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        smp.step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    smp.finish()
    if rank == 0:
        ms = float(t.item()) / args.steps
        print(json.dumps({"slab_bench": {"grid": [H, W], "batch": B, "world": world, "transport": args.transport,
                                         "ms_per_step": ms, "sample_steps_per_s": B / (ms / 1e3),
                                         "pixel_steps_per_s": B * H * W / (ms / 1e3)}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
