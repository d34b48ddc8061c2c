"""The oracle (oracle/guided_sampler_ref.py) against the golden vectors written from the unmodified reference."""
import numpy as np
import torch

from conftest import load_golden, net_from_golden
from oracle import guided_sampler_ref as R


def test_laplacian_and_adjoint():
    g = load_golden("laplacian.npz")
    for tag in "abcd":
        u, dx = g[f"{tag}_u"], float(g[f"{tag}_dx"])
        lap = R.laplacian(torch.from_numpy(u), dx).numpy()
        np.testing.assert_allclose(lap, g[f"{tag}_lap"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(R.laplacian_numpy(u, dx), g[f"{tag}_lap"], rtol=1e-12, atol=1e-12 * np.abs(g[f"{tag}_lap"]).max())
        adj = R.laplacian_adjoint_numpy(g[f"{tag}_gout"], dx)
        np.testing.assert_allclose(adj, g[f"{tag}_adj"], rtol=1e-12, atol=1e-12 * np.abs(g[f"{tag}_adj"]).max())


def test_heat_loss2_and_vjp():
    g = load_golden("pde_losses.npz")
    for tag in ("h1", "h2"):
        u = torch.from_numpy(g[f"{tag}_u"]).requires_grad_()
        d = torch.from_numpy(g[f"{tag}_dudt"]).requires_grad_()
        lab, dx = torch.from_numpy(g[f"{tag}_labels"]), float(g[f"{tag}_dx"])
        loss = R.heat_loss2(u, d, lab, dx)
        gu, gd = torch.autograd.grad(loss, [u, d])
        np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"], rtol=1e-14)
        np.testing.assert_allclose(gu.numpy(), g[f"{tag}_gu"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(gd.numpy(), g[f"{tag}_gdudt"], rtol=1e-12, atol=1e-15)
        # closed-form seed gradient (what the CUDA kernel implements) against reference autograd
        B, _, H, W = g[f"{tag}_u"].shape
        z = np.zeros((B, 0, H, W))
        _, gx, gdn = R.heat_guidance_numpy(g[f"{tag}_u"], g[f"{tag}_dudt"], g[f"{tag}_labels"][:, -1].astype(np.float64), dx,
                                           z, 0.0, np.zeros((H, W)), np.zeros((H, W)), 0, 0.0, 0.0, 1.0)
        np.testing.assert_allclose(gx, g[f"{tag}_gu"], rtol=1e-10, atol=1e-13 * np.abs(g[f"{tag}_gu"]).max())
        np.testing.assert_allclose(gdn, g[f"{tag}_gdudt"], rtol=1e-10, atol=1e-15)


def test_llg_loss2_and_vjp():
    g = load_golden("pde_losses.npz")
    for tag in ("l1", "l2"):
        m = torch.from_numpy(g[f"{tag}_m"]).requires_grad_()
        loss = R.llg_loss2(m, torch.zeros_like(m), None)
        (gm,) = torch.autograd.grad(loss, [m])
        np.testing.assert_allclose(loss.item(), g[f"{tag}_loss"], rtol=1e-14)
        np.testing.assert_allclose(gm.numpy(), g[f"{tag}_gm"], rtol=1e-12, atol=1e-16)


def _joint(gold, C, ch_a, label_dim, loss_fn, loss_kwargs, fd, shape, extra=None):
    torch.set_num_threads(1)
    net = net_from_golden(gold, C, label_dim)
    e = extra or gold
    z = gold["zetas"]
    return R.joint_sample(net, torch.device("cpu"), shape, C, ch_a, loss_fn, loss_kwargs,
                          torch.from_numpy(gold["labels"]), torch.from_numpy(e["obs_a"]), torch.from_numpy(e["obs_u"]),
                          torch.from_numpy(e["mask_a"]), torch.from_numpy(e["mask_u"]), float(z[0]), float(z[1]), float(z[2]),
                          num_steps=int(e["num_steps"]), out_and_grad_fn=fd, latents=torch.from_numpy(e["latents"]))


def test_joint_heat_trajectory():
    gold = load_golden("joint_heat.npz")
    x, losses = _joint(gold, 2, 1, 2, R.heat_loss2, {"dx": float(gold["dx"])}, R.X_and_dXdt_fd, (16, 12))
    np.testing.assert_allclose(losses, gold["losses"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(x.numpy(), gold["x"], rtol=1e-4, atol=1e-5)


def test_joint_heat_empty_mask_branch():
    gold = load_golden("joint_heat.npz")
    e = load_golden("joint_heat_emptymask.npz")
    x, losses = _joint(gold, 2, 1, 2, R.heat_loss2, {"dx": float(gold["dx"])}, R.X_and_dXdt_fd, (16, 12), extra=e)
    assert np.all(losses[:, 1] == 0.0)          # loss_u is the constant-zero branch (sample.py:337-340)
    np.testing.assert_allclose(losses, e["losses"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(x.numpy(), e["x"], rtol=1e-4, atol=1e-5)


def test_joint_llg_trajectory():
    gold = load_golden("joint_llg.npz")
    x, losses = _joint(gold, 6, 3, 4, R.llg_loss2, {}, R.X_and_dXdt_dummy, (16, 8))
    np.testing.assert_allclose(losses, gold["losses"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(x.numpy(), gold["x"], rtol=1e-4, atol=1e-5)


def test_unconditional_trajectory():
    gold, u = load_golden("joint_heat.npz"), load_golden("unconditional_heat.npz")
    torch.set_num_threads(1)
    net = net_from_golden(gold, 2, 2)
    x = R.unconditional_sample(net, torch.device("cpu"), (16, 12), 2, labels=torch.from_numpy(u["labels"]),
                               num_steps=int(u["num_steps"]), latents=torch.from_numpy(u["latents"]))
    np.testing.assert_allclose(x.numpy(), u["x"], rtol=1e-5, atol=1e-6)


def test_weight_switch_and_schedule():
    # i <= 0.8 N in Python floats (sample.py:348): N=20 -> steps 17..19 reduced; N=50 -> 41..; N=200 -> 161..
    for N, first in [(20, 17), (50, 41), (200, 161), (12, 10)]:
        reduced = [i for i in range(N) if R.guidance_weights(i, N, 1.0, 1.0, 1.0)[0] != 1.0]
        assert reduced[0] == first and reduced[-1] == N - 1
    s = R.karras_sigmas(18, 0.002, 80.0, 7.0, "cpu")
    assert s.dtype == torch.float64 and s.shape == (19,) and s[-1] == 0 and abs(s[0].item() - 80.0) < 1e-12
    assert abs(s[17].item() - 0.002) < 1e-15
