"""Config 4 -- zeta / num-steps sensitivity sweep (heat 64x64, N in {20, 50, 200} x 8 zeta triples, 4096 samples) on
1..8 GPUs, one rank per GPU under torchrun.  Prints one JSON line with the aggregate sample-steps/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29641 \
        scripts/sweep_check.py [--samples 4096 --chunk 512 --steps 20,50,200 --zetas 8]
A reduced run (the default: 256 samples, chunks of 128, N = 20, 50) finishes in about a minute on one GPU."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dynamical_pde_diffusion_b200 as dp  # noqa: E402
from dynamical_pde_diffusion_b200 import distributed as D, synthetic  # noqa: E402
from dynamical_pde_diffusion_b200.denoiser import build_unet_v2, randomize_zero_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--steps", default="20,50")
    ap.add_argument("--zetas", type=int, default=8)
    args = ap.parse_args()
    rank, world, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    H = W = 64
    torch.manual_seed(1234)
    net = build_unet_v2(2, 2).eval()
    randomize_zero_init(net, seed=99)
    net = net.to(dev)
    prob = synthetic.heat_problem(1, H, W, seed=0)
    steps = tuple(int(s) for s in args.steps.split(","))
    # 8 zeta triples log-spaced around the reference defaults (conf/sampling_conf/heat_logt_joint.yaml:2-8)
    zetas = [(20.0 * 10 ** (k / (args.zetas - 1) * 2 - 1), 0.5 * 10 ** (k / (args.zetas - 1) * 2 - 1), 20.0) for k in range(args.zetas)]
    make = lambda n, N: dp.JointSampler(net, dev, (H, W), 2, n, 1, dp.heat_loss2, {"dx": prob["dx"]}, num_steps=N)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    final, n_done = D.run_sweep(make, prob, zetas, steps, args.samples, args.chunk, seed=0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    tot = torch.tensor([float(n_done), dt], dtype=torch.float64, device=dev)
    if world > 1:
        work = tot[:1].clone()
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
        tmax = tot[1:].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tot = torch.cat([work, tmax])
    if rank == 0:
        print(json.dumps({"sweep": {"grid": [H, W], "samples": args.samples, "chunk": args.chunk, "steps": steps, "zetas": len(zetas),
                                    "world": world, "seconds": round(float(tot[1]), 3),
                                    "sample_steps_per_s": float(tot[0] / tot[1]),
                                    "final_loss_comb": [[round(float(v), 5) for v in row] for row in final[..., 3]]}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
