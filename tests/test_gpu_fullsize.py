"""Parity at BASELINE.json's full sizes.

The closed-form numpy oracle of tests/test_gpu_kernels.py finishes in seconds only on small grids; here the torch
statement of the reference's losses (oracle/guided_sampler_ref.py: heat_loss2 = pde_losses.py:91-94, the m x H_eff
residual = tests/test_llg_pde_loss.py:70-117, obs_losses = sample.py:336-342) runs in fp64 ON THE DEVICE with autograd,
at the shapes the roofline numbers are quoted on:

* config 5: heat, 8 x 2 x 4096^2 (row-marching heat kernels, strips + chunks + a-plane streaming items),
* LLG m x H_eff residual and the soft |m| = 1 loss, 8 x 6 x 2048^2 (row-marching LLG kernels, the default on this size),
* config 2's per-GPU shard, 64 x 2 x 128^2 (narrow-grid layout: one row per warp),

plus the size-independent properties the path offers: the seed gradient is the derivative of the reduced loss
(directional finite difference through the kernels themselves), the kernels do not depend on the chunk / strip layout,
and the streaming update kernels are bit-exact against torch fp64 on 2^28 elements.

Tolerances: sums 1e-11 relative (fp64, different summation order over 1.3e8 terms), seed gradients 2e-6 of the largest
reference entry (they are rounded to fp32 once; north_star: 1e-5).
"""
import pytest
import torch

from oracle import guided_sampler_ref as R

pytestmark = pytest.mark.gpu

W3 = (20.0, 0.5, 20.0)


def _dev():
    return torch.device("cuda:0")


def _need_gib(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip(f"needs {gib} GiB of device memory")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def _torch_guidance(x0, dxdt, obs_a, obs_u, mask_a, mask_u, ch_a, pde_loss, w):
    """loss = w_a L_a + w_u L_u + w_pde L_pde in fp64 with autograd on the device (sample.py:336-346)."""
    x = x0.double().requires_grad_(True)
    d = dxdt.double().requires_grad_(True)
    la, lu = R.obs_losses(x, obs_a.double(), obs_u.double(), mask_a.double(), mask_u.double(), ch_a)
    lp = pde_loss(x[:, ch_a:], d[:, ch_a:])
    total = w[0] * la.sum() + w[1] * lu.sum() + w[2] * lp
    gx, gd = torch.autograd.grad(total, (x, d), allow_unused=True)
    return (float(la.detach().sum()), float(lu.detach().sum()), float(lp.detach())), gx, gd


def _heat_inputs(B, H, W, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = _dev()
    x0 = torch.randn(B, 2, H, W, device=dev, generator=g)
    dxdt = 0.3 * torch.randn(B, 2, H, W, device=dev, generator=g)
    obs_a, obs_u = torch.randn(1, 1, H, W, device=dev, generator=g), torch.randn(1, 1, H, W, device=dev, generator=g)
    mask_a, mask_u = torch.rand(H, W, device=dev, generator=g) < 0.3, torch.rand(H, W, device=dev, generator=g) < 0.1
    alpha = torch.exp(-2.5 + 3 * torch.rand(B, device=dev, generator=g)).double()
    return x0, dxdt, obs_a, obs_u, mask_a, mask_u, alpha


@pytest.mark.parametrize("shape", [(8, 4096, 4096), (64, 128, 128), (2, 4096, 1000)], ids=["config5", "config2_shard", "ragged_strips"])
def test_heat_guidance_full_size_against_the_torch_oracle_on_device(shape):
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    B, H, W = shape
    _need_gib(40 if H * W * B > 1e8 else 2)
    x0, dxdt, obs_a, obs_u, mask_a, mask_u, alpha = _heat_inputs(B, H, W, seed=H + W)
    dx = 1.0 / (H - 1)
    eng = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, _dev(), obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u, sample_coef=alpha, dx=dx)
    g, gd = eng.seed(x0, dxdt, W3, want_dxdt_grad=True)
    labels = alpha[:, None]
    losses, gx_ref, gd_ref = _torch_guidance(x0, dxdt, obs_a, obs_u, mask_a, mask_u, 1, lambda u, du: R.heat_loss2(u, du, labels, dx), W3)
    got = eng.scalars[:3].cpu().tolist()
    for name, a, b in zip(("L_a", "L_u", "L_pde"), got, losses):
        assert abs(a - b) <= 1e-11 * abs(b), (name, a, b)
    assert _rel(g, gx_ref) < 2e-6
    assert _rel(gd, gd_ref) < 2e-6
    assert torch.all(gd[:, :1] == 0)


@pytest.mark.parametrize("K0", [0.0, 4.0e4])
def test_llg_residual_full_size_against_the_torch_oracle_on_device(K0):
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_RESIDUAL

    B, H, W, ch_a = 8, 2048, 2048, 3
    _need_gib(60)
    dev = _dev()
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.randn(B, 6, H, W, device=dev, generator=g)
    m = x0[:, ch_a:]
    x0[:, ch_a:] = m / m.norm(dim=1, keepdim=True) * (1 + 0.05 * torch.randn(B, 1, H, W, device=dev, generator=g))
    dxdt = 0.01 * torch.randn(B, 6, H, W, device=dev, generator=g)
    field = (30 * torch.randn(B, 3, device=dev, generator=g)).double()
    obs_a, obs_u = torch.randn(1, 3, H, W, device=dev, generator=g), torch.randn(1, 3, H, W, device=dev, generator=g)
    mask_a, mask_u = torch.rand(H, W, device=dev, generator=g) < 0.3, torch.rand(H, W, device=dev, generator=g) < 0.2
    dx, w = 500e-9 / 64, (10.0, 0.5, 10.0)
    c, rc = LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8)), R.LLGConstants(K0=K0, easy_axis=(0.6, 0.0, 0.8))
    eng = GuidanceEngine(B, 6, ch_a, H, W, PDE_LLG_RESIDUAL, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u,
                         sample_coef=field / (1000 * c.mu0), dx=dx, llg=c)
    gx, gd = eng.seed(x0, dxdt, w, want_dxdt_grad=True)
    gx2, _ = eng.seed(x0, dxdt, w)                       # without d / d dmdt the VJP runs its three-CTA kernel (TMA-fed lean items,
    assert _rel(gx2, gx) < 1e-6                          # scatter form: the fp64 rounding differs, an fp32 result may move by one ulp)
    from dynamical_pde_diffusion_b200 import _ffi
    s_tma = eng.scalars[:4].clone()
    try:                                                 # reduce pass fed by cp.async (round-robin item order) instead of TMA
        _ffi.check(_ffi.lib().dpde_set_tuning(7, 1))
        eng.reduce(x0, dxdt, w)
        assert torch.allclose(eng.scalars[:4], s_tma, rtol=1e-12, atol=0)     # other item order: the last bits of the fp64 sums may differ
    finally:
        _ffi.check(_ffi.lib().dpde_set_tuning(7, 0))
    losses, gx_ref, gd_ref = _torch_guidance(x0, dxdt, obs_a, obs_u, mask_a, mask_u, ch_a,
                                             lambda mm, dm: R.llg_residual_loss(mm, dm, field, dx, rc), w)
    got = eng.scalars[:3].cpu().tolist()
    for name, a, b in zip(("L_a", "L_u", "L_pde"), got, losses):
        assert abs(a - b) <= 1e-11 * abs(b), (name, a, b)
    assert _rel(gx, gx_ref) < 2e-6
    assert _rel(gd[:, ch_a:], gd_ref[:, ch_a:]) < 2e-6


def test_llg_norm_full_size_against_the_torch_oracle_on_device():
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_LLG_NORM

    B, H, W, ch_a = 8, 2048, 2048, 3
    _need_gib(60)
    dev = _dev()
    g = torch.Generator(device="cuda").manual_seed(6)
    x0 = torch.randn(B, 6, H, W, device=dev, generator=g)
    obs_a, obs_u = torch.randn(1, 3, H, W, device=dev, generator=g), torch.randn(1, 3, H, W, device=dev, generator=g)
    mask_a, mask_u = torch.rand(H, W, device=dev, generator=g) < 0.3, torch.rand(H, W, device=dev, generator=g) < 0.2
    eng = GuidanceEngine(B, 6, ch_a, H, W, PDE_LLG_NORM, dev, obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u)
    gx, _ = eng.seed(x0, None, W3)
    losses, gx_ref, _ = _torch_guidance(x0, torch.zeros_like(x0), obs_a, obs_u, mask_a, mask_u, ch_a,
                                        lambda mm, dm: R.llg_loss2(mm, dm, None), W3)
    got = eng.scalars[:3].cpu().tolist()
    for name, a, b in zip(("L_a", "L_u", "L_pde"), got, losses):
        assert abs(a - b) <= 1e-11 * abs(b), (name, a, b)
    assert _rel(gx, gx_ref) < 2e-6


def test_seed_is_the_derivative_of_the_reduced_loss_at_full_size():
    """Size-independent property: (loss(x + e v) - loss(x - e v)) / 2e = <seed, v>, both sides from the kernels."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    B, H, W = 8, 4096, 4096
    _need_gib(20)
    x0, dxdt, obs_a, obs_u, mask_a, mask_u, alpha = _heat_inputs(B, H, W, seed=3)
    dx = 1.0 / (H - 1)
    eng = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, _dev(), obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u, sample_coef=alpha, dx=dx)
    g, gd = eng.seed(x0, dxdt, W3, want_dxdt_grad=True)
    gen = torch.Generator(device="cuda").manual_seed(11)
    v = torch.randn(x0.shape, device=_dev(), generator=gen)
    vd = torch.randn(x0.shape, device=_dev(), generator=gen)
    # the perturbation is 2.4e-4 of the fields (second-order term of the square roots ~ 1e-7 relative); what is applied is
    # e v rounded to the fp32 grid of x0, and the analytic side uses exactly that
    e = 2.0 ** -12
    xp, xm = x0 + e * v, x0 - e * v
    dp, dm = dxdt + e * vd, dxdt - e * vd
    dvx, dvd = (xp.double() - xm.double()) / 2, (dp.double() - dm.double()) / 2       # the perturbation actually applied

    def total(x, d):
        eng.reduce(x, d, W3)
        s = eng.scalars[:3].cpu().tolist()
        return W3[0] * s[0] + W3[1] * s[1] + W3[2] * s[2]

    fd = (total(xp, dp) - total(xm, dm)) / 2
    an = float((g.double() * dvx).sum() + (gd.double() * dvd).sum())
    assert abs(fd - an) <= 1e-5 * abs(an), (fd, an)


@pytest.mark.parametrize("tune", [{2: 32}, {2: 128}, {0: 1}, {0: 2}])
def test_layout_knobs_do_not_change_full_size_results(tune):
    from dynamical_pde_diffusion_b200 import GuidanceEngine, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT

    B, H, W = 4, 2048, 4096
    _need_gib(8)
    x0, dxdt, obs_a, obs_u, mask_a, mask_u, alpha = _heat_inputs(B, H, W, seed=9)

    def run():
        eng = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, _dev(), obs_a=obs_a, mask_a=mask_a, obs_u=obs_u, mask_u=mask_u, sample_coef=alpha,
                             dx=1.0 / (H - 1))
        g, gd = eng.seed(x0, dxdt, W3, want_dxdt_grad=True)
        return eng.scalars[:4].clone(), g, gd

    s0, g0, gd0 = run()
    try:
        for k, v in tune.items():
            _ffi.check(_ffi.lib().dpde_set_tuning(k, v))
        s1, g1, gd1 = run()
    finally:
        for k in tune:
            _ffi.check(_ffi.lib().dpde_set_tuning(k, 0))
    assert torch.allclose(s0, s1, rtol=1e-12, atol=0)          # the order of the partial sums follows the layout
    assert torch.equal(g0, g1) or _rel(g1, g0) < 1e-7          # per-pixel arithmetic does not (c_p carries the sums' last bits)
    assert torch.equal(gd0, gd1) or _rel(gd1, gd0) < 1e-7


def test_update_kernels_bit_exact_on_2_to_28_elements():
    """Same fp64 operations as torch (sample.py:327-334, 353-356), bit for bit, at 2^28 elements (config 5's state is 2^28)."""
    from dynamical_pde_diffusion_b200 import _ffi

    n = 1 << 28
    _need_gib(30)
    dev, s = _dev(), torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(2)
    lat = torch.randn(n, device=dev, dtype=torch.float64, generator=g)
    x64, x32 = torch.empty(n, device=dev, dtype=torch.float64), torch.empty(n, device=dev)
    _ffi.call("dpde_sampler_init", lat.data_ptr(), 80.0, x64.data_ptr(), x32.data_ptr(), n, s)
    assert torch.equal(x64, lat * 80.0) and torch.equal(x32, (lat * 80.0).float())
    del lat
    x0c, x0n = torch.randn(n, device=dev, generator=g), torch.randn(n, device=dev, generator=g)
    geu, gcur = torch.randn(n, device=dev, generator=g), torch.randn(n, device=dev, generator=g)
    s_cur, s_next = 3.7123456789, 2.2987654321
    sc, sn = torch.tensor(s_cur, dtype=torch.float64, device=dev), torch.tensor(s_next, dtype=torch.float64, device=dev)
    d_cur = (x64 - x0c.double()) / sc
    x_eu = x64 + (sn - sc) * d_cur
    o64, o32 = torch.empty_like(x64), torch.empty_like(x32)
    _ffi.call("dpde_euler_predict", x64.data_ptr(), x0c.data_ptr(), s_cur, s_next, o32.data_ptr(), n, s)
    assert torch.equal(o32, x_eu.float())
    _ffi.call("dpde_euler_predict_bwd", geu.data_ptr(), s_cur, s_next, o32.data_ptr(), n, s)
    assert torch.equal(o32, (-(((sn - sc) * geu.double()) / sc)).float())
    d_prime = (x_eu - x0n.double()) / sn
    del x_eu
    x_heun = x64 + (sn - sc) * (0.5 * d_cur + 0.5 * d_prime)
    del d_prime, d_cur
    grad = (geu.double() + ((sn - sc) * geu.double()) / sc) + gcur.double()
    _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0c.data_ptr(), x0n.data_ptr(), geu.data_ptr(), gcur.data_ptr(), s_cur, s_next,
              o64.data_ptr(), o32.data_ptr(), n, s)
    want = x_heun - grad
    assert torch.equal(o64, want) and torch.equal(o32, want.float())
