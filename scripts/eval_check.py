"""evaluation.test_loop on several GPUs (one rank per GPU under torchrun): observations are dealt round-robin, the error maps
are all-gathered once; rank 0 compares them with a single-process loop over the same observations (seeded latents).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29661 scripts/eval_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dynamical_pde_diffusion_b200 as dp  # noqa: E402
from dynamical_pde_diffusion_b200 import distributed as D, evaluation as E  # noqa: E402
from conftest import load_golden, net_from_golden  # noqa: E402


def main():
    rank, world, local = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.use_deterministic_algorithms(True, warn_only=True)
    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2, device=dev)
    H, W, B, N, n_obs = 16, 12, 3, 4, 5
    g = torch.Generator().manual_seed(5)
    loader = [{"A": torch.randn(1, 1, H, W, generator=g), "U": torch.randn(1, 1, H, W, generator=g),
               "labels": torch.tensor([[0.2 + 0.05 * i, 0.3]])} for i in range(n_obs)]
    mask_a, mask_u = E.get_masks((H, W), 0.2, 0.2, 0.05, 0.05, generator=g)
    smp = dp.JointSampler(net, dev, (H, W), 2, B, 1, dp.heat_loss2, {"dx": 1.0 / (H - 1)}, num_steps=N)
    res = E.test_loop(smp, loader, 20.0, 0.5, 20.0, mask_a=mask_a, mask_u=mask_u, seed=100, tf32=False,
                      keep_on_device=True)      # uses the default group
    assert res["MAE"].shape == (n_obs, 2, H, W) and res["denom_range"].shape == (n_obs, 2)
    if rank == 0:
        errs = {}
        for i, batch in enumerate(loader):
            lat = D.full_latents(B, 2, (H, W), 100 + i)
            x, _ = smp.sample(batch["labels"].expand(B, -1), batch["A"], batch["U"], mask_a, mask_u, 20.0, 0.5, 20.0, latents=lat)
            mae, d_abs, d_range, std = E.observation_metrics(torch.cat([batch["A"], batch["U"]], dim=1), x)
            for name, ref in (("MAE", mae), ("denom_abs", d_abs[0]), ("denom_range", d_range), ("std", std)):
                e = float(np.abs(res[name][i] - ref.numpy()).max() / max(float(ref.abs().max()), 1e-30))
                errs[name] = max(errs.get(name, 0.0), e)
        assert all(v < 1e-5 for v in errs.values()), errs
        print("eval_check ok: world", world, errs)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
