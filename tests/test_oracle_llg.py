"""LLG m x H_eff residual: oracle restatement vs the reference's per-sample algebra, and closed-form VJP vs autograd."""
import math

import numpy as np
import pytest
import torch

from oracle import guided_sampler_ref as R
from oracle.ref_import import import_reference, reference_available


def _inputs(B=2, H=12, W=9, seed=3):
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(B, 3, H, W, generator=g, dtype=torch.float64)
    m = m / m.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, H, W, generator=g, dtype=torch.float64))
    dmdt = 0.01 * torch.randn(B, 3, H, W, generator=g, dtype=torch.float64)
    field = 30 * torch.randn(B, 3, generator=g, dtype=torch.float64)
    return m, dmdt, field, 500e-9 / 64


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_residual_field_follows_reference_option1():
    """Re-run tests/test_llg_pde_loss.py:70-117 (minus MagTense demag) sample by sample with the reference laplacian."""
    S, _, _ = import_reference()
    m, dmdt, field, dx = _inputs()
    c = R.LLGConstants()
    r = R.llg_residual_field(m, dmdt, field, dx, c)
    for b in range(m.shape[0]):
        h_ext = field[b].view(3, 1, 1) / (1000 * c.mu0)
        h_exch = (2 * c.A0 / (c.mu0 * c.Ms)) * torch.squeeze(S.laplacian(m[b].unsqueeze(1), dx), dim=1)
        h_eff = h_ext + torch.zeros_like(m[b]) + h_exch
        mxH = torch.cross(m[b], h_eff, dim=0)
        rhs = -c.gamma * mxH - c.alpha * torch.cross(m[b], mxH, dim=0)
        ref = dmdt[b] - rhs * c.t_per_step * 1
        torch.testing.assert_close(r[b], ref, rtol=1e-13, atol=1e-18)


@pytest.mark.parametrize("K0", [0.0, 5e4])
def test_residual_closed_form_vjp_matches_autograd(K0):
    m, dmdt, field, dx = _inputs()
    c = R.LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8))
    mt, dt = m.clone().requires_grad_(), dmdt.clone().requires_grad_()
    labels = torch.cat([torch.zeros(m.shape[0], 1, dtype=torch.float64), field], 1)
    loss = R.llg_residual_loss(mt, dt, labels, dx, c)
    gm, gd = torch.autograd.grad(loss, [mt, dt])
    l2, gm2, gd2 = R.llg_residual_guidance_numpy(m.numpy(), dmdt.numpy(), field.numpy(), dx, c)
    assert math.isclose(l2, loss.item(), rel_tol=1e-13)
    np.testing.assert_allclose(gm2, gm.numpy(), rtol=1e-9, atol=1e-12 * np.abs(gm.numpy()).max())
    np.testing.assert_allclose(gd2, gd.numpy(), rtol=1e-10, atol=1e-18)


def test_norm_closed_form_vjp_matches_autograd():
    m, _, _, _ = _inputs()
    mt = m.clone().requires_grad_()
    loss = R.llg_loss2(mt, None, None)
    (gm,) = torch.autograd.grad(loss, [mt])
    l2, g2 = R.llg_norm_guidance_numpy(m.numpy())
    assert math.isclose(l2, loss.item(), rel_tol=1e-13)
    np.testing.assert_allclose(g2, gm.numpy(), rtol=1e-11, atol=1e-18)
