"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container (``/root/reference`` present):  ``python -m oracle.make_golden``.
The reference ships no golden vectors, so parity is pinned by executing its
own functions -- ``laplacian``, ``heat_loss2``, ``llg_loss2`` (with torch autograd for the seed
gradients), ``EDMWrapper(EDMUNet)`` and the full ``JointSampler.sample`` -- on small seeded inputs and
storing inputs and outputs.  The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def golden_laplacian(S):
    g = torch.Generator().manual_seed(11)
    out = {}
    for tag, (n, h, w), dx in [("a", (3, 9, 7), 0.125), ("b", (2, 16, 32), 1.0 / 15), ("c", (1, 2, 2), 0.5),
                               ("d", (2, 40, 5), 7.8125e-9)]:
        u = torch.randn(n, 1, h, w, generator=g, dtype=torch.float64)
        out[f"{tag}_u"], out[f"{tag}_dx"], out[f"{tag}_lap"] = _np(u), np.float64(dx), _np(S.laplacian(u, dx))
        # adjoint by autograd through the reference op
        u2 = u.clone().requires_grad_()
        gout = torch.randn(n, 1, h, w, generator=g, dtype=torch.float64)
        (S.laplacian(u2, dx) * gout).sum().backward()
        out[f"{tag}_gout"], out[f"{tag}_adj"] = _np(gout), _np(u2.grad)
    np.savez_compressed(os.path.join(OUT, "laplacian.npz"), **out)


def golden_pde_losses(PL):
    g = torch.Generator().manual_seed(12)
    out = {}
    # heat_loss2 with gradients w.r.t. u and dudt; inputs are fp32-representable like the net output
    for tag, (b, h, w) in [("h1", (3, 12, 10)), ("h2", (2, 32, 32))]:
        u = torch.randn(b, 1, h, w, generator=g).double().requires_grad_()
        dudt = (0.3 * torch.randn(b, 1, h, w, generator=g)).double().requires_grad_()
        labels = torch.stack([torch.rand(b, generator=g), torch.exp(-2.5 + 3 * torch.rand(b, generator=g))], 1).float()
        dx = 1.0 / (h - 1)
        loss = PL.heat_loss2(u, dudt, labels, dx)
        gu, gd = torch.autograd.grad(loss, [u, dudt])
        out.update({f"{tag}_u": _np(u), f"{tag}_dudt": _np(dudt), f"{tag}_labels": _np(labels), f"{tag}_dx": np.float64(dx),
                    f"{tag}_loss": _np(loss), f"{tag}_gu": _np(gu), f"{tag}_gdudt": _np(gd)})
    # llg_loss2 (soft unit norm)
    for tag, (b, h, w) in [("l1", (2, 16, 8)), ("l2", (3, 64, 16))]:
        m = torch.randn(b, 3, h, w, generator=g).double().requires_grad_()
        loss = PL.llg_loss2(m, torch.zeros_like(m), None)
        (gm,) = torch.autograd.grad(loss, [m])
        out.update({f"{tag}_m": _np(m), f"{tag}_loss": _np(loss), f"{tag}_gm": _np(gm)})
    np.savez_compressed(os.path.join(OUT, "pde_losses.npz"), **out)


def _tiny_net(M, img_channels, label_dim, seed):
    torch.manual_seed(seed)
    unet = M.EDMUNet(img_channels=img_channels, label_dim=label_dim, base_channels=8, channel_mults=(1, 2),
                     num_res_blocks=1, sigma_emb_dim=8, emb_dim=16)
    net = M.EDMWrapper(unet, sigma_data=0.5)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # re-randomise the zero-initialised convolutions (otherwise D(x) = c_skip x)
        for p in net.parameters():
            if float(p.abs().sum()) == 0.0 and p.ndim == 4:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / np.sqrt(p[0].numel())))
    return net.eval()


def _run_joint(S, net, loss_fn, loss_kwargs, out_and_grad_fn, C, ch_a, shape, labels, obs_a, obs_u, mask_a, mask_u,
               zetas, num_steps, seed):
    B = labels.shape[0]
    torch.manual_seed(seed)
    latents = torch.randn((B, C, *shape), dtype=torch.float64)       # what sample.py:314 will draw
    sampler = S.JointSampler(net=net, device=torch.device("cpu"), sample_shape=shape, num_channels=C, num_samples=B,
                             ch_a=ch_a, loss_fn=loss_fn, loss_kwargs=loss_kwargs, num_steps=num_steps,
                             out_and_grad_fn=out_and_grad_fn)
    torch.manual_seed(seed)
    x, losses = sampler.sample(labels, obs_a, obs_u, mask_a, mask_u, *zetas, return_losses=True)
    return latents, x, losses


def golden_joint(S, PL, M):
    torch.use_deterministic_algorithms(True)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(13)
    # ---- heat: C=2, non-square grid, FD time derivative, weight switch inside the run (N=12: steps 10,11 reduced)
    H, W, B, N = 16, 12, 3, 12
    net = _tiny_net(M, 2, 2, seed=21)
    labels = torch.stack([0.5 * torch.rand(B, generator=g), torch.exp(-2.5 + 3 * torch.rand(B, generator=g))], 1).float()
    obs_a = torch.randn(1, 1, H, W, generator=g)
    obs_u = torch.randn(1, 1, H, W, generator=g)
    mask_a = torch.rand(H, W, generator=g) < 0.3
    mask_u = torch.rand(H, W, generator=g) < 0.1
    dx = 1.0 / (H - 1)
    lat, x, losses = _run_joint(S, net, PL.heat_loss2, {"dx": dx}, S.X_and_dXdt_fd, 2, 1, (H, W), labels, obs_a, obs_u,
                                mask_a, mask_u, (20.0, 0.5, 20.0), N, seed=5)
    out = {"latents": _np(lat), "x": _np(x), "losses": losses, "labels": _np(labels), "obs_a": _np(obs_a),
           "obs_u": _np(obs_u), "mask_a": _np(mask_a), "mask_u": _np(mask_u), "dx": np.float64(dx),
           "zetas": np.array([20.0, 0.5, 20.0]), "num_steps": np.int64(N)}
    # one denoiser evaluation, to pin the state-dict-compatible network
    xin = torch.randn(B, 2, H, W, generator=g)
    sig = torch.tensor([0.3, 2.0, 40.0])
    out["net_in"], out["net_sigma"], out["net_out"] = _np(xin), _np(sig), _np(net(xin, sig, labels))
    out.update({f"net/{k}": _np(v) for k, v in net.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "joint_heat.npz"), **out)

    # ---- heat with an empty u-mask (constant-zero branch, sample.py:337-340) and (ch,H,W) masks, obs (ch,H,W)
    mask_u0 = torch.zeros(1, H, W, dtype=torch.bool)
    lat, x, losses = _run_joint(S, net, PL.heat_loss2, {"dx": dx}, S.X_and_dXdt_fd, 2, 1, (H, W), labels, obs_a[0], obs_u[0],
                                mask_a[None], mask_u0, (20.0, 0.5, 20.0), 6, seed=6)
    np.savez_compressed(os.path.join(OUT, "joint_heat_emptymask.npz"), latents=_np(lat), x=_np(x), losses=losses,
                        mask_a=_np(mask_a[None]), mask_u=_np(mask_u0), obs_a=_np(obs_a[0]), obs_u=_np(obs_u[0]),
                        num_steps=np.int64(6))

    # ---- LLG joint: C=6 (a=3, m=3), llg_loss2 + dummy time derivative (test2.py:91-93), 16x8 grid
    H, W, B, N = 16, 8, 2, 8
    net = _tiny_net(M, 6, 4, seed=31)
    labels = torch.cat([torch.rand(B, 1, generator=g), 30 * torch.randn(B, 3, generator=g)], 1).float()
    obs_a = torch.randn(1, 3, H, W, generator=g)
    obs_u = torch.randn(1, 3, H, W, generator=g)
    mask_a = torch.rand(H, W, generator=g) < 0.3
    mask_u = torch.rand(H, W, generator=g) < 0.2
    lat, x, losses = _run_joint(S, net, PL.llg_loss2, {}, S.X_and_dXdt_dummy, 6, 3, (H, W), labels, obs_a, obs_u,
                                mask_a, mask_u, (10.0, 0.5, 10.0), N, seed=7)
    out = {"latents": _np(lat), "x": _np(x), "losses": losses, "labels": _np(labels), "obs_a": _np(obs_a),
           "obs_u": _np(obs_u), "mask_a": _np(mask_a), "mask_u": _np(mask_u),
           "zetas": np.array([10.0, 0.5, 10.0]), "num_steps": np.int64(N)}
    out.update({f"net/{k}": _np(v) for k, v in net.state_dict().items()})
    np.savez_compressed(os.path.join(OUT, "joint_llg.npz"), **out)


def golden_unconditional(S, M):
    """UnconditionalSampler.sample (sample.py:145-239) with the denoiser of joint_heat.npz (same seed -> same weights)."""
    torch.use_deterministic_algorithms(True)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(17)
    H, W, B, N = 16, 12, 3, 9
    net = _tiny_net(M, 2, 2, seed=21)
    labels = torch.stack([0.5 * torch.rand(B, generator=g), torch.exp(-2.5 + 3 * torch.rand(B, generator=g))], 1).float()
    torch.manual_seed(8)
    latents = torch.randn((B, 2, H, W), dtype=torch.float64)          # what sample.py:222 will draw
    sampler = S.UnconditionalSampler(net=net, device=torch.device("cpu"), sample_shape=(H, W), num_channels=2, num_samples=B,
                                     num_steps=N)
    torch.manual_seed(8)
    x = sampler.sample(labels=labels)
    np.savez_compressed(os.path.join(OUT, "unconditional_heat.npz"), latents=_np(latents), x=_np(x), labels=_np(labels),
                        num_steps=np.int64(N))


def golden_edm_heat_loss(M):
    """EDMHeatLoss.__call__ (models/loss.py:127-171) with the denoiser of joint_heat.npz; the two internal torch.randn
    draws are reproduced from the same seed and stored, so a test can feed them back (``noise=``)."""
    torch.use_deterministic_algorithms(True)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(23)
    H, W, B = 16, 12, 3
    net = _tiny_net(M, 2, 2, seed=21)
    x = torch.randn(B, 2, H, W, generator=g)
    labels = torch.stack([0.5 * torch.rand(B, generator=g), torch.exp(-2.5 + 3 * torch.rand(B, generator=g))], 1).float()
    dx = 1.0 / (H - 1)
    out = {"x": _np(x), "labels": _np(labels), "dx": np.float64(dx)}
    w0 = next(p for p in net.parameters() if p.ndim == 4)
    for tag, kw in (("me_mean", dict(residual_estimation="ME", reduce_method="mean")),
                    ("me_sum", dict(residual_estimation="ME", reduce_method="sum")),
                    ("se_mean", dict(residual_estimation="SE", reduce_method="mean"))):
        seed = 40 + len(tag)
        torch.manual_seed(seed)
        rnd = torch.randn([B, 1, 1, 1])
        eps = torch.randn_like(x)
        torch.manual_seed(seed)
        loss = M.EDMHeatLoss(dx, pde_loss_coeff=0.37, **kw)(net, x, labels)
        (gw,) = torch.autograd.grad(loss.mean(), [w0])
        out.update({f"{tag}_rnd": _np(rnd), f"{tag}_eps": _np(eps), f"{tag}_loss": _np(loss), f"{tag}_gw0": _np(gw)})
    np.savez_compressed(os.path.join(OUT, "edm_heat_loss.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    S, PL, M = import_reference()
    if "--only-unconditional" in sys.argv:
        golden_unconditional(S, M)
        return
    if "--only-edm-heat-loss" in sys.argv:
        golden_edm_heat_loss(M)
        return
    golden_laplacian(S)
    golden_pde_losses(PL)
    golden_joint(S, PL, M)
    golden_unconditional(S, M)
    golden_edm_heat_loss(M)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
