"""GPU parity of the CUDA kernels (called through the C ABI) against the oracle.  Run with ``-m gpu`` on a B200.

Tolerance (north_star): <= 1e-5 relative in fp32, measured against the largest reference magnitude of the field
(``_close``); integer / mask / branch handling must be exact.  fp64 outputs are held to 1e-11.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import guided_sampler_ref as R

pytestmark = pytest.mark.gpu

RTOL32, RTOL64 = 1e-5, 1e-11


def _close(got, ref, rtol, what=""):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, np.float64)
    ref = ref.detach().double().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-300)
    err = float(np.abs(got - ref).max()) / scale
    assert err <= rtol, f"{what}: max error {err:.3e} of field scale > {rtol}"


def _dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------------------
# laplacian and its transpose
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_laplacian_golden(dtype):
    import dynamical_pde_diffusion_b200 as dp

    g = load_golden("laplacian.npz")
    for tag in "abcd":
        u = torch.from_numpy(g[f"{tag}_u"]).to(_dev(), dtype).requires_grad_()
        dx = float(g[f"{tag}_dx"])
        out = dp.laplacian(u, dx)
        rt = RTOL64 if dtype == torch.float64 else RTOL32
        ref_lap = g[f"{tag}_lap"] if dtype == torch.float64 else R.laplacian_numpy(u.detach().double().cpu().numpy(), dx)
        _close(out, ref_lap, rt, f"lap {tag}")
        (out * torch.from_numpy(g[f"{tag}_gout"]).to(_dev(), dtype)).sum().backward()
        _close(u.grad, g[f"{tag}_adj"], rt if dtype == torch.float64 else 2e-5, f"lap adjoint {tag}")


def test_laplacian_rejects_multichannel_and_cpu():
    import dynamical_pde_diffusion_b200 as dp

    with pytest.raises(RuntimeError):
        dp.laplacian(torch.zeros(1, 3, 8, 8, device=_dev()), 0.1)   # reference conv weight is (1,1,3,3)
    with pytest.raises(RuntimeError):
        dp.laplacian(torch.zeros(1, 1, 8, 8), 0.1)                   # no CPU path


# ------------------------------------------------------------------------------------------------------------
# Level-1 losses against the golden vectors (reference functions + reference autograd)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_heat_loss2_golden(dtype):
    import dynamical_pde_diffusion_b200 as dp

    g = load_golden("pde_losses.npz")
    rt = RTOL64 if dtype == torch.float64 else RTOL32
    for tag in ("h1", "h2"):
        u = torch.from_numpy(g[f"{tag}_u"]).to(_dev(), dtype).requires_grad_()
        d = torch.from_numpy(g[f"{tag}_dudt"]).to(_dev(), dtype).requires_grad_()
        lab = torch.from_numpy(g[f"{tag}_labels"]).to(_dev())
        loss = dp.heat_loss2(u, d, lab, np.float64(g[f"{tag}_dx"]))       # numpy dx, as test2.py:84-87 passes it
        assert loss.dim() == 0 and loss.dtype == dtype
        gu, gd = torch.autograd.grad(3.0 * loss, [u, d])
        _close(loss, g[f"{tag}_loss"], rt, "heat loss")
        _close(gu, 3.0 * g[f"{tag}_gu"], rt, "heat d/du")
        _close(gd, 3.0 * g[f"{tag}_gdudt"], rt, "heat d/ddudt")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_llg_loss2_golden(dtype):
    import dynamical_pde_diffusion_b200 as dp

    g = load_golden("pde_losses.npz")
    rt = RTOL64 if dtype == torch.float64 else RTOL32
    for tag in ("l1", "l2"):
        m = torch.from_numpy(g[f"{tag}_m"]).to(_dev(), dtype).requires_grad_()
        loss = dp.llg_loss2(m, torch.zeros_like(m), None)
        (gm,) = torch.autograd.grad(loss, [m])
        _close(loss, g[f"{tag}_loss"], rt, "llg norm loss")
        _close(gm, g[f"{tag}_gm"], rt, "llg norm grad")


def test_level1_accepts_channel_slice_views():
    """The unmodified sampler passes x_N[:, ch_a:] of an fp64 NCHW tensor (sample.py:345-346)."""
    import dynamical_pde_diffusion_b200 as dp

    gen = torch.Generator().manual_seed(2)
    x = torch.randn(3, 2, 20, 24, generator=gen).double().to(_dev()).requires_grad_()
    dxdt = torch.randn(3, 2, 20, 24, generator=gen).double().to(_dev())
    lab = torch.rand(3, 2, generator=gen).to(_dev())
    loss = dp.heat_loss2(x[:, 1:], dxdt[:, 1:], lab, 0.05)
    (gx,) = torch.autograd.grad(loss, [x])
    xr = x.detach().clone().requires_grad_()
    lr = R.heat_loss2(xr[:, 1:], dxdt[:, 1:], lab, 0.05)
    (gr,) = torch.autograd.grad(lr, [xr])
    _close(loss, lr, RTOL64, "loss on view")
    _close(gx, gr, RTOL64, "grad on view")
    assert torch.all(gx[:, 0] == 0)


# ------------------------------------------------------------------------------------------------------------
# fused guidance (three losses + seed gradient) against the closed-form numpy oracle
# ------------------------------------------------------------------------------------------------------------
def _heat_case(B, ch_a, cu, H, W, seed, mask_mode="hw", dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    C_ = ch_a + cu
    x0 = torch.randn(B, C_, H, W, generator=gen).to(dtype)
    dxdt = (0.5 * torch.randn(B, C_, H, W, generator=gen)).to(dtype)
    labels = torch.stack([torch.rand(B, generator=gen), torch.exp(-2.5 + 3 * torch.rand(B, generator=gen))], 1).float()
    if mask_mode == "hw":            # bool (H,W) masks, obs (1,ch,H,W): test2.py:49, model_testing.py:192
        mask_a, mask_u = torch.rand(H, W, generator=gen) < 0.3, torch.rand(H, W, generator=gen) < 0.1
        obs_a, obs_u = torch.randn(1, ch_a, H, W, generator=gen), torch.randn(1, cu, H, W, generator=gen)
    elif mask_mode == "chw":         # (ch,H,W) bool masks and obs
        mask_a, mask_u = torch.rand(ch_a, H, W, generator=gen) < 0.3, torch.rand(cu, H, W, generator=gen) < 0.2
        obs_a, obs_u = torch.randn(ch_a, H, W, generator=gen), torch.randn(cu, H, W, generator=gen)
    else:                            # per-sample float masks with non-binary weights, fp64 observations
        mask_a, mask_u = torch.rand(B, ch_a, H, W, generator=gen), torch.rand(B, cu, H, W, generator=gen)
        obs_a = torch.randn(B, ch_a, H, W, generator=gen, dtype=torch.float64)
        obs_u = torch.randn(B, cu, H, W, generator=gen, dtype=torch.float64)
    return x0, dxdt, labels, obs_a, obs_u, mask_a, mask_u


@pytest.fixture(params=["march", "generic"])
def kernel_path(request):
    """Run a test once with the register row-marching kernels enabled (taken when the layout is eligible) and once
    with the generic tile kernels forced."""
    from dynamical_pde_diffusion_b200 import _ffi

    old = _ffi.lib().dpde_set_fast_path(1 if request.param == "march" else 0)
    yield request.param
    _ffi.lib().dpde_set_fast_path(old)


@pytest.mark.parametrize("shape", [(3, 1, 1, 12, 10), (2, 1, 1, 2, 2), (2, 1, 1, 64, 64), (1, 1, 1, 37, 130), (2, 1, 1, 130, 17),
                                   (4, 1, 1, 128, 128), (1, 2, 2, 33, 65), (2, 0, 1, 16, 16), (1, 1, 1, 40, 256),
                                   (1, 1, 1, 24, 520), (3, 1, 1, 9, 16), (2, 1, 1, 5, 8), (1, 1, 1, 3, 4), (2, 2, 2, 70, 132)])
@pytest.mark.parametrize("mask_mode", ["hw", "chw", "full"])
def test_heat_guidance_matches_closed_form(shape, mask_mode, kernel_path):
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    B, ch_a, cu, H, W = shape
    if ch_a == 0 and mask_mode != "hw":
        pytest.skip("no a-channels")
    x0, dxdt, labels, obs_a, obs_u, mask_a, mask_u = _heat_case(B, ch_a, cu, H, W, seed=H * 1000 + W, mask_mode=mask_mode)
    dx, w = 1.0 / (H - 1), (20.0, 0.5, 20.0)
    dev = _dev()
    eng = GuidanceEngine(B, ch_a + cu, ch_a, H, W, PDE_HEAT, dev, obs_a=obs_a.to(dev) if ch_a else None,
                         mask_a=mask_a.to(dev) if ch_a else None, obs_u=obs_u.to(dev), mask_u=mask_u.to(dev),
                         sample_coef=labels[:, -1].double().to(dev), dx=dx)
    trace = torch.zeros(4, device=dev)
    g, gd = eng.seed(x0.to(dev), dxdt.to(dev), w, trace_row=trace, want_dxdt_grad=True)
    if ch_a == 0:
        obs_a, mask_a = np.zeros((B, 0, H, W)), np.zeros((H, W))
    losses, g_ref, gd_ref = R.heat_guidance_numpy(x0.double().numpy(), dxdt.double().numpy(), labels[:, -1].double().numpy(), dx,
                                                  np.asarray(obs_a, np.float64), obs_u.double().numpy(), np.asarray(mask_a, np.float64),
                                                  mask_u.double().numpy(), ch_a, *w)
    _close(eng.scalars[:4], np.array(losses), 1e-12, "losses")
    _close(trace, np.array(losses, np.float32), 1e-7, "trace row")
    _close(g, g_ref, RTOL32, "seed gradient")
    _close(gd[:, ch_a:], gd_ref, RTOL32, "d/d dudt")
    assert torch.all(gd[:, :ch_a] == 0)


@pytest.mark.parametrize("tune", [{0: 1}, {0: 2}, {2: 8}, {2: 128}, {3: 1}, {4: 1}, {3: 1, 4: 1, 0: 2}])
def test_tuning_knobs_never_change_results(tune):
    """dpde_set_tuning selects strip layouts / chunk lengths / a-plane pairing of the marching kernels: speed only."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    dev, w = _dev(), (20.0, 0.5, 20.0)
    for (B, ch_a, cu, H, W) in [(2, 1, 1, 70, 260), (3, 1, 1, 33, 128), (1, 2, 2, 40, 520)]:
        x0, dxdt, labels, obs_a, obs_u, mask_a, mask_u = _heat_case(B, ch_a, cu, H, W, seed=H + W)
        def run():
            eng = GuidanceEngine(B, ch_a + cu, ch_a, H, W, PDE_HEAT, dev, obs_a=obs_a.to(dev), mask_a=mask_a.to(dev), obs_u=obs_u.to(dev),
                                 mask_u=mask_u.to(dev), sample_coef=labels[:, -1].double().to(dev), dx=1.0 / (H - 1))
            g, gd = eng.seed(x0.to(dev), dxdt.to(dev), w, want_dxdt_grad=True)
            return eng.scalars[:4].clone(), g, gd
        s0, g0, gd0 = run()
        try:
            for k, v in tune.items():
                _ffi.check(_ffi.lib().dpde_set_tuning(k, v))
            s1, g1, gd1 = run()
        finally:
            for k in tune:
                _ffi.check(_ffi.lib().dpde_set_tuning(k, 0))
        _close(s1, s0, 1e-13, "losses under tuning")          # partial sums are grouped differently: summation order only
        _close(g1, g0, 2e-7, "seed under tuning")
        _close(gd1, gd0, 2e-7, "d/d dudt under tuning")
    assert _ffi.lib().dpde_set_tuning(99, 0) != 0


def test_heat_guidance_fp64_fields_and_autograd_cross_check():
    """fp64 fields (what the unmodified sampler holds) and an independent check against torch autograd on the device."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    B, H, W, dev = 3, 40, 24, _dev()
    x0, dxdt, labels, obs_a, obs_u, mask_a, mask_u = _heat_case(B, 1, 1, H, W, seed=9, dtype=torch.float64)
    x0, dxdt, labels = x0.to(dev), dxdt.to(dev), labels.to(dev)
    obs_a, obs_u, mask_a, mask_u = obs_a.to(dev), obs_u.to(dev), mask_a.to(dev), mask_u.to(dev)
    w, dx = (3.0, 0.7, 11.0), 0.02
    eng = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u,
                         sample_coef=labels[:, -1].double(), dx=dx)
    g, _ = eng.seed(x0, dxdt, w)
    assert g.dtype == torch.float64
    xr = x0.clone().requires_grad_()
    la, lu = R.obs_losses(xr, obs_a.double(), obs_u.double(), mask_a.double(), mask_u.double(), 1)
    lp = R.heat_loss2(xr[:, 1:], dxdt[:, 1:], labels, dx)
    (gr,) = torch.autograd.grad(w[0] * la + w[1] * lu + w[2] * lp, [xr])
    _close(g, gr, RTOL64, "fp64 seed vs autograd")
    _close(eng.scalars[:3], torch.stack([la.reshape(()), lu.reshape(()), lp.reshape(())]), 1e-12, "losses")


def test_empty_mask_branches_and_nan_semantics(kernel_path):
    """mask.sum() == 0 -> the loss is the constant 0 with no gradient (sample.py:337-342);
    a non-empty mask with zero residual -> NaN gradient, as sqrt'(0) gives in the reference."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    B, H, W, dev = 2, 16, 16, _dev()
    x0, dxdt, labels, obs_a, obs_u, mask_a, mask_u = _heat_case(B, 1, 1, H, W, seed=4)
    empty = torch.zeros(H, W, dtype=torch.bool)
    eng = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, dev, obs_a=obs_a.to(dev), mask_a=empty.to(dev), obs_u=obs_u.to(dev),
                         mask_u=empty.to(dev), sample_coef=labels[:, -1].double().to(dev), dx=0.1)
    g, _ = eng.seed(x0.to(dev), dxdt.to(dev), (20.0, 0.5, 20.0))
    assert eng.desc.has_a == 0 and eng.desc.has_u == 0
    assert float(eng.scalars[0]) == 0.0 and float(eng.scalars[1]) == 0.0
    assert torch.all(g[:, 0] == 0) and torch.isfinite(g).all()
    # zero residual under a non-empty mask
    eng2 = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, dev, obs_a=x0[:, :1].to(dev), mask_a=mask_a.to(dev), obs_u=obs_u.to(dev),
                          mask_u=mask_u.to(dev), sample_coef=labels[:, -1].double().to(dev), dx=0.1)
    g2, _ = eng2.seed(x0.to(dev), dxdt.to(dev), (20.0, 0.5, 20.0))
    xr = x0.double().requires_grad_()
    la, lu = R.obs_losses(xr, x0[:, :1].double(), obs_u.double(), mask_a.double(), mask_u.double(), 1)
    (gr,) = torch.autograd.grad(20.0 * la + 0.5 * lu, [xr])
    assert torch.equal(torch.isnan(g2[:, 0]).cpu(), torch.isnan(gr[:, 0]))
    assert torch.isfinite(g2[:, 1]).all()


@pytest.mark.parametrize("shape", [(2, 3, 16, 8), (3, 3, 64, 16), (1, 3, 33, 130), (2, 0, 12, 20), (2, 3, 70, 132), (3, 3, 128, 128),
                                   (1, 3, 5, 4100)])
def test_llg_norm_guidance(shape, kernel_path):
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_NORM

    B, ch_a, H, W = shape
    gen = torch.Generator().manual_seed(H + W)
    x0 = torch.randn(B, ch_a + 3, H, W, generator=gen)
    x0[0, ch_a:, 0, 0] = 0.0                      # |m| = 0 pixel: zero gradient, as torch.linalg.norm's backward
    obs_a, obs_u = torch.randn(1, max(ch_a, 1), H, W, generator=gen), torch.randn(1, 3, H, W, generator=gen)
    mask_a, mask_u = torch.rand(H, W, generator=gen) < 0.3, torch.rand(H, W, generator=gen) < 0.2
    dev, w = _dev(), (10.0, 0.5, 10.0)
    eng = GuidanceEngine(B, ch_a + 3, ch_a, H, W, PDE_LLG_NORM, dev, obs_a=obs_a.to(dev) if ch_a else None,
                         mask_a=mask_a.to(dev) if ch_a else None, obs_u=obs_u.to(dev), mask_u=mask_u.to(dev))
    g, _ = eng.seed(x0.to(dev), None, w)
    xr = x0.double().requires_grad_()
    la, lu = R.obs_losses(xr, obs_a.double() if ch_a else torch.zeros(1), obs_u.double(),
                          mask_a.double() if ch_a else torch.zeros(1), mask_u.double(), ch_a)
    lp = R.llg_loss2(xr[:, ch_a:], None, None)
    (gr,) = torch.autograd.grad(w[0] * la + w[1] * lu + w[2] * lp, [xr])
    _close(g, gr, RTOL32, "llg norm seed")
    _close(eng.scalars[2], lp, 1e-12, "llg norm loss")
    assert torch.isfinite(g).all()


@pytest.mark.parametrize("shape", [(2, 3, 12, 9), (2, 3, 64, 16), (1, 3, 130, 40), (2, 0, 20, 70), (2, 3, 16, 8), (2, 0, 12, 20),
                                   (1, 3, 70, 132), (3, 3, 128, 128), (1, 3, 20, 36), (2, 3, 2, 4), (1, 3, 3, 260), (1, 0, 66, 64)])
@pytest.mark.parametrize("K0", [0.0, 5e4])
def test_llg_residual_guidance(shape, K0, kernel_path):
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_RESIDUAL

    B, ch_a, H, W = shape
    gen = torch.Generator().manual_seed(7 * H + W)
    x0 = torch.randn(B, ch_a + 3, H, W, generator=gen)
    m = x0[:, ch_a:]
    x0[:, ch_a:] = m / m.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, H, W, generator=gen))
    dxdt = 0.01 * torch.randn(B, ch_a + 3, H, W, generator=gen)
    field = (30 * torch.randn(B, 3, generator=gen)).double()
    obs_u, mask_u = torch.randn(1, 3, H, W, generator=gen), torch.rand(H, W, generator=gen) < 0.2
    obs_a, mask_a = torch.randn(1, max(ch_a, 1), H, W, generator=gen), torch.rand(H, W, generator=gen) < 0.3
    dev, w, dx = _dev(), (10.0, 0.5, 10.0), 500e-9 / 64
    c = LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8))
    rc = R.LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8))
    eng = GuidanceEngine(B, ch_a + 3, ch_a, H, W, PDE_LLG_RESIDUAL, dev, obs_a=obs_a.to(dev) if ch_a else None,
                         mask_a=mask_a.to(dev) if ch_a else None, obs_u=obs_u.to(dev), mask_u=mask_u.to(dev),
                         sample_coef=(field / (1000 * c.mu0)).to(dev), dx=dx, llg=c)
    g, gd = eng.seed(x0.to(dev), dxdt.to(dev), w, want_dxdt_grad=True)
    loss, gm_ref, gd_ref = R.llg_residual_guidance_numpy(x0[:, ch_a:].double().numpy(), dxdt[:, ch_a:].double().numpy(),
                                                         field.numpy(), dx, rc, w_pde=w[2])
    mu = np.broadcast_to(mask_u.double().numpy(), (B, 3, H, W))
    du = mu * (x0[:, ch_a:].double().numpy() - obs_u.double().numpy())
    gm_ref = gm_ref + w[1] * mu * du / math.sqrt((du ** 2).sum())
    _close(eng.scalars[2], np.array(loss), 1e-12, "llg residual loss")
    _close(g[:, ch_a:], gm_ref, RTOL32, "llg residual seed")
    _close(gd[:, ch_a:], gd_ref, RTOL32, "llg residual d/d dmdt")


@pytest.mark.parametrize("shape", [(2, 3, 40, 132), (1, 3, 70, 260), (2, 0, 36, 128), (1, 3, 33, 520)])
@pytest.mark.parametrize("rows", [0, 8, 6, 14])
@pytest.mark.parametrize("want_d", [False, True])
def test_llg_residual_marching_kernels(shape, rows, want_d):
    """The row-marching LLG kernels (llg_march.cuh; default on large grids only) forced onto small grids by tuning key 6 = 2:
    against the closed form, with chunk lengths that give the reduce pass (8) resp. the VJP (6, 14: R + 2 a multiple of the
    ring depth) interior work items for the lean loop, with and without the d / d dmdt output (which disables the lean VJP
    loop), with K0 != 0, and against the tile kernels (key 6 = 1), the general loop only (key 5 = 1) and the cp.async feed of the
    reduce pass instead of TMA (key 7 = 1: bit-identical)."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_RESIDUAL

    B, ch_a, H, W = shape
    gen = torch.Generator().manual_seed(3 * H + W + rows)
    x0 = torch.randn(B, ch_a + 3, H, W, generator=gen)
    m = x0[:, ch_a:]
    x0[:, ch_a:] = m / m.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, H, W, generator=gen))
    dxdt = 0.01 * torch.randn(B, ch_a + 3, H, W, generator=gen)
    field = (30 * torch.randn(B, 3, generator=gen)).double()
    obs_u, mask_u = torch.randn(1, 3, H, W, generator=gen), torch.rand(3, H, W, generator=gen) < 0.2
    obs_a, mask_a = torch.randn(1, max(ch_a, 1), H, W, generator=gen), torch.rand(H, W, generator=gen) < 0.3
    dev, w, dx = _dev(), (10.0, 0.5, 10.0), 500e-9 / 64
    K0 = 5e4 if rows == 6 else 0.0
    c, rc = LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8)), R.LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8))

    def run():
        eng = GuidanceEngine(B, ch_a + 3, ch_a, H, W, PDE_LLG_RESIDUAL, dev, obs_a=obs_a.to(dev) if ch_a else None,
                             mask_a=mask_a.to(dev) if ch_a else None, obs_u=obs_u.to(dev), mask_u=mask_u.to(dev),
                             sample_coef=(field / (1000 * c.mu0)).to(dev), dx=dx, llg=c)
        g, gd = eng.seed(x0.to(dev), dxdt.to(dev), w, want_dxdt_grad=want_d)
        return eng.scalars[:4].clone(), g, gd

    T = _ffi.lib().dpde_set_tuning
    try:
        _ffi.check(T(2, rows))
        _ffi.check(T(6, 2))
        s_m, g_m, gd_m = run()                    # marching, lean loop where the geometry allows it (TMA-fed; VJP: three-CTA kernel)
        _ffi.check(T(7, 1))
        s_c, g_c, gd_c = run()                    # the same without TMA: cp.async feed in the reduce pass, two-CTA VJP kernel
        _ffi.check(T(7, 0))
        _ffi.check(T(5, 1))
        s_g, g_g, gd_g = run()                    # marching, general loop only
        _ffi.check(T(5, 0))
        _ffi.check(T(6, 1))
        _ffi.check(T(2, 0))
        s_t, g_t, gd_t = run()                    # convert-once tiles
    finally:
        for k in (2, 5, 6, 7):
            _ffi.check(T(k, 0))
    _close(s_c, s_m, 1e-13, "sums, TMA-fed reduce pass (longest-first item order) vs cp.async feed (round-robin order)")
    _close(g_c, g_m, 2e-7, "seed, three-CTA VJP kernel vs two-CTA kernel")    # scatter form: fp64 rounding differs, fp32 results may move by one ulp
    loss, gm_ref, gd_ref = R.llg_residual_guidance_numpy(x0[:, ch_a:].double().numpy(), dxdt[:, ch_a:].double().numpy(),
                                                         field.numpy(), dx, rc, w_pde=w[2])
    mu = np.broadcast_to(mask_u.double().numpy(), (B, 3, H, W))
    du = mu * (x0[:, ch_a:].double().numpy() - obs_u.double().numpy())
    gm_ref = gm_ref + w[1] * mu * du / math.sqrt((du ** 2).sum())
    _close(s_m[2], np.array(loss), 1e-12, "loss (marching)")
    _close(g_m[:, ch_a:], gm_ref, RTOL32, "seed (marching)")
    _close(s_g, s_m, 1e-13, "losses, general vs lean loop")
    _close(g_g, g_m, 2e-7, "seed, general vs lean loop")
    _close(s_t, s_m, 1e-12, "losses, tiles vs marching")
    _close(g_t, g_m, 4e-7, "seed, tiles vs marching")
    if want_d:
        _close(gd_m[:, ch_a:], gd_ref, RTOL32, "d/d dmdt (marching)")
        assert torch.all(gd_m[:, :ch_a] == 0)
        _close(gd_t, gd_m, 4e-7, "d/d dmdt, tiles vs marching")


@pytest.mark.parametrize("kind_name", ["heat", "llg_residual"])
def test_absent_time_derivative_equals_zeros(kind_name, kernel_path):
    """dxdt == NULL (X_and_dXdt_dummy without the zeros tensor) must give what an all-zero dxdt gives, on both kernel paths;
    per-channel (ch, H, W) masks and per-sample observations exercise the stride handling of the fast paths."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT, PDE_LLG_RESIDUAL

    dev, w = _dev(), (3.0, 0.7, 11.0)
    kind, ch_a, cu, dx = (PDE_HEAT, 1, 1, 1 / 39) if kind_name == "heat" else (PDE_LLG_RESIDUAL, 3, 3, 500e-9 / 64)
    B, H, W = 3, 40, 72
    g = torch.Generator().manual_seed(17)
    x0 = torch.randn(B, ch_a + cu, H, W, generator=g).to(dev)
    obs_a, obs_u = torch.randn(B, ch_a, H, W, generator=g).to(dev), torch.randn(B, cu, H, W, generator=g).to(dev)
    mask_a, mask_u = (torch.rand(ch_a, H, W, generator=g) < 0.3).to(dev), (torch.rand(cu, H, W, generator=g) < 0.2).to(dev)
    coef = torch.rand(B, generator=g).double().to(dev) if kind == PDE_HEAT else (1e4 * torch.randn(B, 3, generator=g)).double().to(dev)
    outs = []
    for dxdt in (None, torch.zeros_like(x0)):
        eng = GuidanceEngine(B, ch_a + cu, ch_a, H, W, kind, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u,
                             sample_coef=coef, dx=dx, llg=LLGConstants())
        gx, _ = eng.seed(x0, dxdt, w)
        outs.append((gx.clone(), eng.scalars[:4].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.isfinite(outs[0][0]).all()


def test_level1_llg_residual_matches_oracle_autograd():
    import dynamical_pde_diffusion_b200 as dp

    gen = torch.Generator().manual_seed(1)
    m = torch.randn(2, 3, 24, 16, generator=gen).double().to(_dev()).requires_grad_()
    d = (0.01 * torch.randn(2, 3, 24, 16, generator=gen)).double().to(_dev()).requires_grad_()
    lab = torch.cat([torch.rand(2, 1, generator=gen), 20 * torch.randn(2, 3, generator=gen)], 1).float().to(_dev())
    dx = 500e-9 / 64
    loss = dp.llg_residual_loss(m, d, lab, dx)
    gm, gd = torch.autograd.grad(loss, [m, d])
    mr, dr = m.detach().clone().requires_grad_(), d.detach().clone().requires_grad_()
    lr = R.llg_residual_loss(mr, dr, lab, dx)
    gmr, gdr = torch.autograd.grad(lr, [mr, dr])
    _close(loss, lr, RTOL64, "loss")
    _close(gm, gmr, 1e-9, "d/dm")
    _close(gd, gdr, 1e-9, "d/d dmdt")


# ------------------------------------------------------------------------------------------------------------
# row slabs: the same fields cut into slabs with ghost rows must give the same sums and the same gradient
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind_name", ["heat", "llg_residual", "llg_norm"])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_row_slab_decomposition_equals_whole_grid(kind_name, n_slabs, kernel_path):
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT, PDE_LLG_NORM, PDE_LLG_RESIDUAL

    dev, halo = _dev(), 2
    gen = torch.Generator().manual_seed(5)
    if kind_name == "heat":
        kind, ch_a, cu, H, W, dx = PDE_HEAT, 1, 1, 48, 40, 1 / 47
    else:
        kind, ch_a, cu, H, W, dx = (PDE_LLG_RESIDUAL if kind_name == "llg_residual" else PDE_LLG_NORM), 3, 3, 36, 20, 500e-9 / 64
    B, C_ = 2, ch_a + cu
    x0 = torch.randn(B, C_, H, W, generator=gen).to(dev)
    dxdt = (0.1 * torch.randn(B, C_, H, W, generator=gen)).to(dev)
    obs_a, obs_u = torch.randn(1, ch_a, H, W, generator=gen).to(dev), torch.randn(1, cu, H, W, generator=gen).to(dev)
    mask_a, mask_u = (torch.rand(H, W, generator=gen) < 0.3).to(dev), (torch.rand(H, W, generator=gen) < 0.2).to(dev)
    coef = torch.rand(B, generator=gen).double().to(dev) if kind == PDE_HEAT else (1e4 * torch.randn(B, 3, generator=gen)).double().to(dev)
    if kind == PDE_LLG_NORM:
        coef = None
    w = (5.0, 0.5, 7.0)
    whole = GuidanceEngine(B, C_, ch_a, H, W, kind, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u,
                           sample_coef=coef, dx=dx, llg=LLGConstants())
    g_ref, gd_ref = whole.seed(x0, dxdt, w, want_dxdt_grad=True)

    rows = [H * r // n_slabs for r in range(n_slabs + 1)]
    engines, locals_ = [], []

    def pad(t, r0, r1):  # local buffer = halo ghost rows + owned rows + halo ghost rows (zeros beyond the grid)
        out = torch.zeros(*t.shape[:-2], r1 - r0 + 2 * halo, W, dtype=t.dtype, device=dev)
        lo, hi = max(r0 - halo, 0), min(r1 + halo, H)
        out[..., lo - (r0 - halo): hi - (r0 - halo), :] = t[..., lo:hi, :]
        return out

    total = torch.zeros(3, dtype=torch.float64, device=dev)
    for r in range(n_slabs):
        r0, r1 = rows[r], rows[r + 1]
        e = GuidanceEngine(B, C_, ch_a, r1 - r0 + 2 * halo, W, kind, dev, obs_a=pad(obs_a, r0, r1), mask_a=pad(mask_a, r0, r1),
                           obs_u=pad(obs_u, r0, r1), mask_u=pad(mask_u, r0, r1), sample_coef=coef, dx=dx, llg=LLGConstants(),
                           slab=dict(halo=halo, row0=r0, H_global=H, has_a=True, has_u=True))
        xl, dl = pad(x0, r0, r1), pad(dxdt, r0, r1)
        e.reduce(xl, dl, w, finalize=False)
        total += e.sums
        engines.append(e)
        locals_.append((xl, dl, r0, r1))
    _close(total, whole.sums, 1e-12, "slab sums")
    for e, (xl, dl, r0, r1) in zip(engines, locals_):
        e.sums.copy_(total)
        e.finalize()
        g, gd = e.vjp(xl, dl, w, want_dxdt_grad=True)
        _close(g[..., halo:-halo, :], g_ref[..., r0:r1, :], 2e-6, "slab gradient")
        _close(gd[..., halo:-halo, :], gd_ref[..., r0:r1, :], 2e-6, "slab d/d dxdt")
        assert torch.all(g[..., :halo, :] == 0) and torch.all(g[..., -halo:, :] == 0)   # ghost rows never written


@pytest.mark.parametrize("rows", [8, 14])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_llg_marching_kernels_on_row_slabs(rows, n_slabs):
    """The row-marching LLG kernels on row slabs (ghost rows, global reflection only at the global top / bottom): chunk length 8 gives
    the reduce pass TMA-fed lean items, 14 gives the VJP its three-CTA kernel (TMA-fed lean items, light general items, work queue);
    no d / d dmdt output, which would send the VJP to its two-CTA kernel.  Against the whole grid and against the kernels without TMA."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_RESIDUAL

    dev, halo, B, ch_a, H, W, dx = _dev(), 2, 2, 3, 120, 260, 500e-9 / 64
    gen = torch.Generator().manual_seed(17 + rows)
    x0 = torch.randn(B, 6, H, W, generator=gen).to(dev)
    dxdt = (0.01 * torch.randn(B, 6, H, W, generator=gen)).to(dev)
    obs_a, obs_u = torch.randn(1, 3, H, W, generator=gen).to(dev), torch.randn(1, 3, H, W, generator=gen).to(dev)
    mask_a, mask_u = (torch.rand(H, W, generator=gen) < 0.3).to(dev), (torch.rand(H, W, generator=gen) < 0.2).to(dev)
    coef, w = (1e4 * torch.randn(B, 3, generator=gen)).double().to(dev), (5.0, 0.5, 7.0)
    bounds = [H * r // n_slabs for r in range(n_slabs + 1)]

    def pad(t, r0, r1):
        out = torch.zeros(*t.shape[:-2], r1 - r0 + 2 * halo, W, dtype=t.dtype, device=dev)
        lo, hi = max(r0 - halo, 0), min(r1 + halo, H)
        out[..., lo - (r0 - halo): hi - (r0 - halo), :] = t[..., lo:hi, :]
        return out

    def run():
        whole = GuidanceEngine(B, 6, ch_a, H, W, PDE_LLG_RESIDUAL, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u, sample_coef=coef,
                               dx=dx, llg=LLGConstants())
        g_ref, _ = whole.seed(x0, dxdt, w)
        engines, total = [], torch.zeros(3, dtype=torch.float64, device=dev)
        for r in range(n_slabs):
            r0, r1 = bounds[r], bounds[r + 1]
            e = GuidanceEngine(B, 6, ch_a, r1 - r0 + 2 * halo, W, PDE_LLG_RESIDUAL, dev, obs_a=pad(obs_a, r0, r1), mask_a=pad(mask_a, r0, r1),
                               obs_u=pad(obs_u, r0, r1), mask_u=pad(mask_u, r0, r1), sample_coef=coef, dx=dx, llg=LLGConstants(),
                               slab=dict(halo=halo, row0=r0, H_global=H, has_a=True, has_u=True))
            xl, dl = pad(x0, r0, r1), pad(dxdt, r0, r1)
            e.reduce(xl, dl, w, finalize=False)
            total += e.sums
            engines.append((e, xl, dl, r0, r1))
        _close(total, whole.sums, 1e-12, "slab sums")
        pieces = []
        for e, xl, dl, r0, r1 in engines:
            e.sums.copy_(total)
            e.finalize()
            g, _ = e.vjp(xl, dl, w)
            _close(g[..., halo:-halo, :], g_ref[..., r0:r1, :], 2e-6, "slab gradient")
            assert torch.all(g[..., :halo, :] == 0) and torch.all(g[..., -halo:, :] == 0)      # ghost rows never written
            pieces.append(g[..., halo:-halo, :])
        return whole.sums.clone(), g_ref, torch.cat(pieces, dim=-2)

    T = _ffi.lib().dpde_set_tuning
    try:
        _ffi.check(T(6, 2))
        _ffi.check(T(2, rows))
        s1, gw1, gs1 = run()                       # TMA forms
        _ffi.check(T(7, 1))
        s0, gw0, gs0 = run()                       # cp.async feed, two-CTA VJP kernel
    finally:
        for k in (2, 6, 7):
            _ffi.check(T(k, 0))
    _close(s1, s0, 1e-13, "sums with / without TMA")
    _close(gw1, gw0, 2e-7, "whole-grid seed with / without TMA")
    _close(gs1, gs0, 2e-7, "slab seeds with / without TMA")


@pytest.mark.parametrize("variant", ["no_dxdt", "no_obs_u", "no_a_planes", "batch1_ragged"])
def test_llg_marching_tma_forms_match_the_fallbacks_on_every_template_variant(variant):
    """TMA-fed reduce pass / three-CTA VJP kernel against the cp.async reduce pass / two-CTA VJP kernel (tuning key 7) for the template
    variants the other tests do not reach: no time derivative, empty observation mask, no a-planes, batch 1 on a ragged width."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_RESIDUAL

    dev = _dev()
    B, ch_a, H, W = {"no_dxdt": (2, 3, 64, 260), "no_obs_u": (2, 3, 64, 260), "no_a_planes": (3, 0, 48, 388), "batch1_ragged": (1, 3, 80, 452)}[variant]
    gen = torch.Generator().manual_seed(len(variant))
    x0 = torch.randn(B, ch_a + 3, H, W, generator=gen).to(dev)
    dxdt = None if variant == "no_dxdt" else (0.01 * torch.randn(B, ch_a + 3, H, W, generator=gen)).to(dev)
    obs_u = torch.randn(1, 3, H, W, generator=gen).to(dev)
    mask_u = (torch.rand(H, W, generator=gen) < (0.0 if variant == "no_obs_u" else 0.2)).to(dev)
    obs_a = torch.randn(1, 3, H, W, generator=gen).to(dev) if ch_a else None
    mask_a = (torch.rand(H, W, generator=gen) < 0.3).to(dev) if ch_a else None
    coef = (1e4 * torch.randn(B, 3, generator=gen)).double().to(dev)
    T = _ffi.lib().dpde_set_tuning

    def run():
        eng = GuidanceEngine(B, ch_a + 3, ch_a, H, W, PDE_LLG_RESIDUAL, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u,
                             sample_coef=coef, dx=500e-9 / 64, llg=LLGConstants())
        g, _ = eng.seed(x0, dxdt, (10.0, 0.5, 10.0))
        return eng.scalars[:4].clone(), g

    for rows in (8, 14):                                  # 8: lean items in the reduce pass, 14: in the VJP
        try:
            _ffi.check(T(6, 2))
            _ffi.check(T(2, rows))
            s1, g1 = run()
            _ffi.check(T(7, 1))
            s0, g0 = run()
        finally:
            for k in (2, 6, 7):
                _ffi.check(T(k, 0))
        _close(s1, s0, 1e-13, f"sums ({variant}, rows {rows})")
        _close(g1, g0, 2e-7, f"seed ({variant}, rows {rows})")


def test_llg_vjp_work_queue_counters_wrap_around():
    """The three-CTA LLG VJP kernel takes one zeroed work-queue counter per launch from a ring of 256 device slots: 600 launches wrap
    the ring twice, on two streams; every launch must reproduce the first result bit for bit."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_RESIDUAL

    dev, B, H, W = _dev(), 2, 64, 260
    gen = torch.Generator().manual_seed(23)
    x0, dxdt = torch.randn(B, 6, H, W, generator=gen).to(dev), (0.01 * torch.randn(B, 6, H, W, generator=gen)).to(dev)
    obs_a, obs_u = torch.randn(1, 3, H, W, generator=gen).to(dev), torch.randn(1, 3, H, W, generator=gen).to(dev)
    mask = (torch.rand(H, W, generator=gen) < 0.2).to(dev)
    coef = (1e4 * torch.randn(B, 3, generator=gen)).double().to(dev)
    T = _ffi.lib().dpde_set_tuning
    try:
        _ffi.check(T(6, 2))
        _ffi.check(T(2, 14))
        eng = GuidanceEngine(B, 6, 3, H, W, PDE_LLG_RESIDUAL, dev, obs_a=obs_a, mask_a=mask, obs_u=obs_u, mask_u=mask, sample_coef=coef,
                             dx=500e-9 / 64, llg=LLGConstants())
        first, _ = eng.seed(x0, dxdt, (10.0, 0.5, 10.0))
        first = first.clone()
        side = torch.cuda.Stream()
        for i in range(600):
            if i % 2:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    g, _ = eng.vjp(x0, dxdt, (10.0, 0.5, 10.0))
                torch.cuda.current_stream().wait_stream(side)
            else:
                g, _ = eng.vjp(x0, dxdt, (10.0, 0.5, 10.0))
            if i % 50 == 49:
                assert torch.equal(g, first), f"launch {i} differs"
        torch.cuda.synchronize()
        assert torch.equal(g, first)
    finally:
        for k in (2, 6):
            _ffi.check(T(k, 0))


def test_halo_pack_unpack_roundtrip():
    import ctypes as C
    from dynamical_pde_diffusion_b200 import _ffi

    dev, planes, H, W, halo = _dev(), 6, 20, 24, 2
    f = torch.randn(planes, H, W, device=dev, dtype=torch.float64)
    up, down = torch.empty(planes, halo, W, device=dev, dtype=torch.float64), torch.empty(planes, halo, W, device=dev, dtype=torch.float64)
    s = torch.cuda.current_stream().cuda_stream
    _ffi.call("dpde_halo_pack", f.data_ptr(), _ffi.F64, planes, H, W, halo, up.data_ptr(), down.data_ptr(), s)
    assert torch.equal(up, f[:, halo:2 * halo]) and torch.equal(down, f[:, H - 2 * halo:H - halo])
    g = torch.zeros_like(f)
    _ffi.call("dpde_halo_unpack", g.data_ptr(), _ffi.F64, planes, H, W, halo, up.data_ptr(), None, s)
    assert torch.equal(g[:, :halo], up) and torch.all(g[:, halo:] == 0)
    _ffi.call("dpde_halo_unpack", g.data_ptr(), _ffi.F64, planes, H, W, halo, None, down.data_ptr(), s)
    assert torch.equal(g[:, H - halo:], down)


# ------------------------------------------------------------------------------------------------------------
# Euler / Heun / update kernels: same fp64 operations as torch, bit for bit
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 + 5, 2 * 2 * 64 * 64])
def test_update_kernels_bit_exact(n):
    from dynamical_pde_diffusion_b200 import _ffi

    dev = _dev()
    gen = torch.Generator().manual_seed(n)
    s = torch.cuda.current_stream().cuda_stream
    lat = torch.randn(n, generator=gen, dtype=torch.float64).to(dev)
    s_cur, s_next = 3.7123456789, 2.2987654321
    x64, x32 = torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
    _ffi.call("dpde_sampler_init", lat.data_ptr(), 80.0, x64.data_ptr(), x32.data_ptr(), n, s)
    assert torch.equal(x64, lat * 80.0) and torch.equal(x32, (lat * 80.0).float())

    x0c = torch.randn(n, generator=gen).to(dev)
    x0n = torch.randn(n, generator=gen).to(dev)
    geu = torch.randn(n, generator=gen).to(dev)
    gcur = torch.randn(n, generator=gen).to(dev)
    sc, sn = torch.tensor(s_cur, dtype=torch.float64, device=dev), torch.tensor(s_next, dtype=torch.float64, device=dev)
    d_cur = (x64 - x0c.double()) / sc                                            # sample.py:327-334, same op order
    x_eu = x64 + (sn - sc) * d_cur
    out32 = torch.empty_like(x32)
    _ffi.call("dpde_euler_predict", x64.data_ptr(), x0c.data_ptr(), s_cur, s_next, out32.data_ptr(), n, s)
    assert torch.equal(out32, x_eu.float())

    seed = torch.empty_like(x32)
    _ffi.call("dpde_euler_predict_bwd", geu.data_ptr(), s_cur, s_next, seed.data_ptr(), n, s)
    assert torch.equal(seed, (-(((sn - sc) * geu.double()) / sc)).float())

    d_prime = (x_eu - x0n.double()) / sn
    x_heun = x64 + (sn - sc) * (0.5 * d_cur + 0.5 * d_prime)
    grad = (geu.double() + ((sn - sc) * geu.double()) / sc) + gcur.double()
    o64, o32 = torch.empty_like(x64), torch.empty_like(x32)
    _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0c.data_ptr(), x0n.data_ptr(), geu.data_ptr(), gcur.data_ptr(),
              s_cur, s_next, o64.data_ptr(), o32.data_ptr(), n, s)
    assert torch.equal(o64, x_heun - grad) and torch.equal(o32, (x_heun - grad).float())
    # last step: Euler only, sigma_next = 0, gradient through the single denoiser evaluation
    zero = torch.zeros((), dtype=torch.float64, device=dev)
    _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0c.data_ptr(), None, None, gcur.data_ptr(), s_cur, 0.0,
              o64.data_ptr(), o32.data_ptr(), n, s)
    assert torch.equal(o64, (x64 + (zero - sc) * d_cur) - gcur.double())


def test_c_abi_reports_errors():
    import ctypes as C
    from dynamical_pde_diffusion_b200 import _ffi

    L = _ffi.lib()
    assert L.dpde_euler_predict(None, None, 1.0, 0.5, None, 4, None) == -1
    assert b"null" in L.dpde_last_error()
    d = _ffi.GuidanceDesc()
    d.B, d.C, d.ch_a, d.H, d.W, d.pde_kind = 1, 2, 1, 1, 8, _ffi.PDE_HEAT     # H = 1 cannot be reflect-padded
    assert L.dpde_guidance_vjp(C.byref(d), None, None, None, None, None) == -1
    with pytest.raises(_ffi.DpdeError):
        _ffi.call("dpde_laplacian", None, None, 0, 1, 8, 8, 64, 0.1, 0, None)
