"""Per-step parity probe (GPU): feed OUR state into the oracle's guided_step and compare one step at a time;
also measures run-to-run nondeterminism of both.  Diagnostic tool, not a test."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dynamical_pde_diffusion_b200 as dp  # noqa: E402
from conftest import load_golden, net_from_golden  # noqa: E402
from oracle import guided_sampler_ref as R  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
if "--det" in sys.argv:
    torch.backends.cudnn.deterministic = True
    torch.use_deterministic_algorithms(True, warn_only=True)
dev = torch.device("cuda:0")
gold = load_golden("joint_heat.npz")
net = net_from_golden(gold, 2, 2, device=dev)
z = gold["zetas"]
N = int(gold["num_steps"])
kw = {"dx": float(gold["dx"])}
T = lambda k: torch.from_numpy(gold[k])


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


smp = dp.JointSampler(net, dev, (16, 12), 2, 3, 1, dp.heat_loss2, kw, num_steps=N)
run = smp.begin(T("labels"), T("obs_a"), T("obs_u"), T("mask_a"), T("mask_u"), float(z[0]), float(z[1]), float(z[2]), latents=T("latents"))
sig = torch.tensor(run["sigmas"], dtype=torch.float64, device=dev)
oa, ou = T("obs_a").to(dev).double(), T("obs_u").to(dev).double()
ma, mu = T("mask_a").to(dev).double(), T("mask_u").to(dev).double()
lab = T("labels").to(dev)
for i in range(N):
    x_in = run["x64"].clone()
    xo, row, it = R.guided_step(net, x_in, i, sig, lab, oa, ou, ma, mu, 1, R.heat_loss2, kw, float(z[0]), float(z[1]), float(z[2]), N,
                                return_internals=True)
    xo2, row2 = R.guided_step(net, x_in, i, sig, lab, oa, ou, ma, mu, 1, R.heat_loss2, kw, float(z[0]), float(z[1]), float(z[2]), N)
    smp.step()
    ours = run["x64"]
    tr = run["trace"][i].cpu().numpy()
    print(f"step {i:2d} sigma {run['sigmas'][i]:9.4f}: x_next rel err {rel(ours, xo):.2e}  oracle-rerun {rel(xo2, xo):.2e}  "
          f"loss rel err {np.abs(tr - np.float32(row)).max() / np.abs(np.float32(row)).max():.2e}  |grad|/|x| {float(it['grad_x'].abs().max() / xo.abs().max()):.2e}")
