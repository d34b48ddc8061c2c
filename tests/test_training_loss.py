"""Training-time physics loss: EDMHeatLoss (models/loss.py:41-171).

CPU: the oracle restatement against golden vectors written from the UNMODIFIED reference class (the two internal
torch.randn draws are stored with them).  GPU: the CUDA per-sample residual kernels against the torch expression, and
the drop-in class against the oracle on the same device and against the golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_golden, net_from_golden
from oracle import guided_sampler_ref as R

CASES = (("me_mean", dict(residual_estimation="ME", reduce_method="mean")),
         ("me_sum", dict(residual_estimation="ME", reduce_method="sum")),
         ("se_mean", dict(residual_estimation="SE", reduce_method="mean")))


def _first_conv_weight(net):
    return next(p for p in net.parameters() if p.ndim == 4)


def test_oracle_edm_heat_loss_matches_the_reference():
    gold, e = load_golden("joint_heat.npz"), load_golden("edm_heat_loss.npz")
    torch.set_num_threads(1)
    net = net_from_golden(gold, 2, 2)
    x, labels, dx = torch.from_numpy(e["x"]), torch.from_numpy(e["labels"]), float(e["dx"])
    for tag, kw in CASES:
        noise = (torch.from_numpy(e[f"{tag}_rnd"]), torch.from_numpy(e[f"{tag}_eps"]))
        loss = R.edm_heat_loss(net, x, labels, dx, noise, pde_loss_coeff=0.37, **kw)
        assert loss.shape == (3, 1, 1, 3)          # the reference broadcasts (B,) against sigma (B,1,1,1): loss.py:146
        np.testing.assert_allclose(loss.detach().numpy(), e[f"{tag}_loss"], rtol=2e-5)
        (gw,) = torch.autograd.grad(loss.mean(), [_first_conv_weight(net)])
        np.testing.assert_allclose(gw.numpy(), e[f"{tag}_gw0"], rtol=1e-3, atol=1e-4 * np.abs(e[f"{tag}_gw0"]).max())


def test_training_ops_refuse_cpu():
    from dynamical_pde_diffusion_b200 import training as T

    u = torch.randn(2, 1, 8, 8)
    with pytest.raises(RuntimeError):
        T.heat_residual_sq(u, u, torch.ones(2), 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        T.EDMHeatLoss(0.1)(None, torch.randn(2, 2, 8, 8), torch.ones(2, 2))


def _rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else b
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 1, 16, 12), (2, 1, 64, 64), (1, 2, 33, 130), (5, 1, 2, 2), (2, 1, 128, 256)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_heat_residual_sq_matches_torch(shape, dtype):
    from dynamical_pde_diffusion_b200 import training as T

    dev = torch.device("cuda:0")
    B, Cu, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    full = torch.randn(B, Cu + 1, H, W, generator=g).to(dtype).to(dev)
    u = full[:, 1:].requires_grad_()                     # a channel-slice view, as x_0star[:, ch_a:] (loss.py:143)
    d = (0.3 * torch.randn(B, Cu, H, W, generator=g)).to(dtype).to(dev).requires_grad_()
    alpha = torch.exp(-2.5 + 3 * torch.rand(B, generator=g)).to(dev)
    dx = 1.0 / (max(H, 2) - 1)
    out = T.heat_residual_sq(u, d, alpha, dx)
    w = torch.rand(B, generator=g).to(dev).to(dtype)
    gu, gd = torch.autograd.grad((out * w).sum(), [u, d])
    # reference expression in fp64 (planes one by one: the reference's laplacian takes a single channel)
    u64, d64 = u.detach().double().requires_grad_(), d.detach().double().requires_grad_()
    lap = torch.cat([R.laplacian(u64[:, c:c + 1], dx) for c in range(Cu)], dim=1)
    ref = ((d64 - alpha.double().view(-1, 1, 1, 1) * lap) ** 2).sum(dim=(1, 2, 3))
    gur, gdr = torch.autograd.grad((ref * w.double()).sum(), [u64, d64])
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    assert out.shape == (B,) and out.dtype == dtype
    assert _rel(out.cpu(), ref.detach().cpu()) < tol
    assert _rel(gu.cpu(), gur.cpu()) < tol and _rel(gd.cpu(), gdr.cpu()) < tol


@pytest.mark.gpu
def test_edm_heat_loss_against_oracle_and_golden():
    import dynamical_pde_diffusion_b200.training as T

    dev = torch.device("cuda:0")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    try:
        gold, e = load_golden("joint_heat.npz"), load_golden("edm_heat_loss.npz")
        net = net_from_golden(gold, 2, 2, device=dev)
        w0 = _first_conv_weight(net)
        x, labels, dx = torch.from_numpy(e["x"]).to(dev), torch.from_numpy(e["labels"]).to(dev), float(e["dx"])
        for tag, kw in CASES:
            noise = (torch.from_numpy(e[f"{tag}_rnd"]).to(dev), torch.from_numpy(e[f"{tag}_eps"]).to(dev))
            loss = T.EDMHeatLoss(dx, pde_loss_coeff=0.37, **kw)(net, x, labels, noise=noise)
            (gw,) = torch.autograd.grad(loss.mean(), [w0])
            ref = R.edm_heat_loss(net, x, labels, dx, noise, pde_loss_coeff=0.37, **kw)
            (gwr,) = torch.autograd.grad(ref.mean(), [w0])
            assert loss.shape == ref.shape == (3, 1, 1, 3)
            # same device, same denoiser kernels: what differs is the fp32 conv2d Laplacian of the reference vs our
            # fp64 stencil rounded once (fp32 rounding, no amplification: the time derivative is shared)
            assert _rel(loss.detach().cpu(), ref.detach().cpu()) < 1e-5, (tag, _rel(loss.detach().cpu(), ref.detach().cpu()))
            assert _rel(gw.cpu(), gwr.cpu()) < 1e-4, (tag, _rel(gw.cpu(), gwr.cpu()))
            # CPU-generated golden: different convolution kernels, amplified by the finite-difference time derivative
            assert _rel(loss.detach().cpu(), e[f"{tag}_loss"]) < 5e-2, (tag, _rel(loss.detach().cpu(), e[f"{tag}_loss"]))
        # the unseeded path draws its own noise
        out = T.EDMHeatLoss(dx)(net, x, labels)
        assert out.shape == (3, 1, 1, 3) and torch.isfinite(out).all()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old
