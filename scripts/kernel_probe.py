"""Launch each hand-written kernel a few times on a large grid (default: config 5's 8x2x4096^2) -- the target of the
ncu captures under profiles/.  Kernel-only, no denoiser.  Usage: python scripts/kernel_probe.py [B H W] [--llg]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants, _ffi  # noqa: E402
from dynamical_pde_diffusion_b200._ffi import PDE_HEAT, PDE_LLG_NORM, PDE_LLG_RESIDUAL  # noqa: E402

nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
B, H, W = nums if len(nums) == 3 else (8, 4096, 4096)
llg = "--llg" in sys.argv
reps = 10
for a in sys.argv[1:]:                      # --tune=key:value[,key:value...]  (dpde_set_tuning)
    if a.startswith("--tune="):
        for kv in a[7:].split(","):
            k, v = kv.split(":")
            _ffi.check(_ffi.lib().dpde_set_tuning(int(k), int(v)))
if "--generic" in sys.argv:
    _ffi.lib().dpde_set_fast_path(0)
dev = torch.device("cuda:0")
C_, ch_a = (6, 3) if llg else (2, 1)
if "--uonly" in sys.argv:
    C_, ch_a = C_ - ch_a, 0
cu = C_ - ch_a
s = torch.cuda.current_stream().cuda_stream
x0 = torch.randn(B, C_, H, W, device=dev)
dxdt = torch.randn(B, C_, H, W, device=dev)
mask = torch.rand(H, W, device=dev) < 0.2
obs_a, obs_u = torch.randn(1, max(ch_a, 1), H, W, device=dev), torch.randn(1, cu, H, W, device=dev)
w = (20.0, 0.5, 20.0)


loop = 0
for a in sys.argv[1:]:
    if a.startswith("--loop="):
        loop = int(a[7:])
    if a.startswith("--reps="):
        reps = int(a[7:])


def timed(label, nbytes, fn):
    fn()
    torch.cuda.synchronize()
    if loop:                                 # back-to-back launches: steady-state (L2-warm) cost per launch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(loop):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / loop
        print(f"{label:34s} {ms * 1e3:8.2f} us/launch (x{loop} back to back)  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic)")
        return
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{label:34s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic)")


if "--train" in sys.argv:                # per-sample residual of the training loss (models/loss.py:143), u / dudt (B,1,H,W)
    from dynamical_pde_diffusion_b200 import training as T
    u = x0[:, -1:].detach().requires_grad_()
    d = dxdt[:, -1:].detach().requires_grad_()
    alpha = torch.rand(B, device=dev)
    px = B * H * W
    hold = {}
    def fwd():
        hold["out"] = T.heat_residual_sq(u, d, alpha, 1.0 / (H - 1))
    timed("heat_residual_sq", 8 * px, fwd)
    up = torch.ones(B, device=dev)
    def bwd():
        hold["g"] = torch.autograd.grad(hold["out"], [u, d], grad_outputs=up, retain_graph=True)
    timed("heat_residual_sq_vjp", 16 * px, bwd)
    sys.exit(0)

kinds = [("llg_residual", PDE_LLG_RESIDUAL), ("llg_norm", PDE_LLG_NORM)] if llg else [("heat", PDE_HEAT)]
for name, kind in kinds:
    coef = None
    if kind == PDE_HEAT:
        coef = torch.rand(B, device=dev).double()
    elif kind == PDE_LLG_RESIDUAL:
        coef = (1e4 * torch.randn(B, 3, device=dev)).double()
    eng = GuidanceEngine(B, C_, ch_a, H, W, kind, dev, obs_a=obs_a if ch_a else None, mask_a=mask if ch_a else None, obs_u=obs_u, mask_u=mask, sample_coef=coef,
                         dx=1.0 / (H - 1) if kind == PDE_HEAT else 500e-9 / 64, llg=LLGConstants())
    d = None if kind == PDE_LLG_NORM else dxdt
    nd = 0 if d is None else cu
    px = B * H * W
    timed(f"guidance_reduce[{name}]", 4 * px * (C_ + nd), lambda: eng.reduce(x0, d, w))
    hold = {}
    def vjp():
        hold["g"] = None
        hold["g"] = eng.vjp(x0, d, w)
    timed(f"guidance_vjp[{name}]", 4 * px * (2 * C_ + nd), vjp)
    del eng, hold

if not llg and "--guidance-only" not in sys.argv:
    n = x0.numel()
    x64 = torch.randn(B, C_, H, W, device=dev, dtype=torch.float64)
    o64, o32 = torch.empty_like(x64), torch.empty_like(x0)
    timed("euler_predict", 16 * n, lambda: _ffi.call("dpde_euler_predict", x64.data_ptr(), x0.data_ptr(), 3.0, 2.0, o32.data_ptr(), n, s))
    timed("euler_predict_bwd", 8 * n, lambda: _ffi.call("dpde_euler_predict_bwd", x0.data_ptr(), 3.0, 2.0, o32.data_ptr(), n, s))
    g2 = torch.randn_like(x0)
    timed("heun_guided_update", 36 * n, lambda: _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0.data_ptr(), dxdt.data_ptr(),
                                                          g2.data_ptr(), o32.data_ptr(), 3.0, 2.0, o64.data_ptr(), o32.data_ptr(), n, s))
    timed("sampler_init", 20 * n, lambda: _ffi.call("dpde_sampler_init", x64.data_ptr(), 80.0, o64.data_ptr(), o32.data_ptr(), n, s))
