"""GPU parity of the drop-in JointSampler: golden trajectories of the unmodified reference (generated on CPU) and
the oracle run on the same device with the same denoiser and latents.  Run with ``-m gpu`` on a B200."""
import numpy as np
import pytest
import torch

from conftest import load_golden, net_from_golden
from oracle import guided_sampler_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _ieee_fp32():
    """Parity runs use IEEE fp32, deterministic cuDNN algorithms.

    TF32 (the reference's sampling_context, sample.py:626-630) is a throughput setting exercised by bench.py.
    Determinism matters: with cuDNN's default (atomics-based) backward kernels the ORACLE ITSELF differs run to run
    by 1e-6..1e-4 per step on this network (scripts/step_parity_probe.py; the finite-difference time derivative
    with eps = 1e-5 in fp32 amplifies one-ulp differences), with deterministic algorithms our sampler and the
    oracle agree bit for bit at every step."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark,
           torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.use_deterministic_algorithms(True, warn_only=True)
    yield
    torch.use_deterministic_algorithms(False)
    (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark,
     torch.backends.cudnn.deterministic) = old


def _dev():
    return torch.device("cuda:0")


def _rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))


def _ours(gold, e, C_, ch_a, label_dim, loss_fn, loss_kwargs, provider, shape, **kw):
    import dynamical_pde_diffusion_b200 as dp

    net = net_from_golden(gold, C_, label_dim, device=_dev())
    z = gold["zetas"]
    smp = dp.JointSampler(net, _dev(), shape, C_, int(gold["labels"].shape[0]), ch_a, loss_fn, loss_kwargs,
                          num_steps=int(e["num_steps"]), out_and_grad_fn=provider)
    return smp.sample(torch.from_numpy(gold["labels"]), torch.from_numpy(e["obs_a"]), torch.from_numpy(e["obs_u"]),
                      torch.from_numpy(e["mask_a"]), torch.from_numpy(e["mask_u"]), float(z[0]), float(z[1]), float(z[2]),
                      return_losses=True, latents=torch.from_numpy(e["latents"]), **kw)


def _oracle_on_device(gold, e, C_, ch_a, label_dim, loss_fn, loss_kwargs, provider, shape):
    net = net_from_golden(gold, C_, label_dim, device=_dev())
    z = gold["zetas"]
    return R.joint_sample(net, _dev(), shape, C_, ch_a, loss_fn, loss_kwargs, torch.from_numpy(gold["labels"]),
                          torch.from_numpy(e["obs_a"]), torch.from_numpy(e["obs_u"]), torch.from_numpy(e["mask_a"]),
                          torch.from_numpy(e["mask_u"]), float(z[0]), float(z[1]), float(z[2]), num_steps=int(e["num_steps"]),
                          out_and_grad_fn=provider, latents=torch.from_numpy(e["latents"]))


# cross-device tolerance: the golden files were produced by the reference on CPU (MKL/oneDNN convolutions); the
# denoiser here runs on cuDNN, so ~1e-6 differences per evaluation are amplified through the guided steps.
XDEV = 5e-3
# same-device tolerance (oracle and ours share the denoiser kernels): north_star's 1e-5
SAME = 1e-5


def test_heat_golden_and_same_device_oracle():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    kw = {"dx": float(gold["dx"])}
    x, losses = _ours(gold, gold, 2, 1, 2, dp.heat_loss2, kw, dp.X_and_dXdt_fd, (16, 12))
    assert isinstance(x, torch.Tensor) and x.device.type == "cpu" and x.dtype == torch.float32 and x.shape == (3, 2, 16, 12)
    assert isinstance(losses, np.ndarray) and losses.shape == (12, 4) and losses.dtype == np.float32
    assert _rel(losses, gold["losses"]) < XDEV and _rel(x.numpy(), gold["x"]) < XDEV
    xo, lo = _oracle_on_device(gold, gold, 2, 1, 2, R.heat_loss2, kw, R.X_and_dXdt_fd, (16, 12))
    assert _rel(losses, lo) < SAME, _rel(losses, lo)
    assert _rel(x.numpy(), xo.numpy()) < SAME, _rel(x.numpy(), xo.numpy())


def test_unconditional_sampler_golden_and_same_device_oracle():
    """UnconditionalSampler (sample.py:145-239) on the fused init / predictor / update kernels."""
    import dynamical_pde_diffusion_b200 as dp

    gold, u = load_golden("joint_heat.npz"), load_golden("unconditional_heat.npz")
    net = net_from_golden(gold, 2, 2, device=_dev())
    labels, lat, N = torch.from_numpy(u["labels"]), torch.from_numpy(u["latents"]), int(u["num_steps"])
    smp = dp.UnconditionalSampler(net, _dev(), (16, 12), 2, 3, num_steps=N)
    x = smp.sample(labels=labels, latents=lat)
    assert isinstance(x, torch.Tensor) and x.device.type == "cpu" and x.dtype == torch.float32 and x.shape == (3, 2, 16, 12)
    assert _rel(x.numpy(), u["x"]) < XDEV
    xo = R.unconditional_sample(net, _dev(), (16, 12), 2, labels=labels, num_steps=N, latents=lat)
    assert torch.equal(x, xo), _rel(x.numpy(), xo.numpy())      # same denoiser kernels + bit-exact update kernels
    # default latents are the first RNG draw (sample.py:222)
    torch.manual_seed(3)
    x1 = smp.sample(labels=labels)
    torch.manual_seed(3)
    lat1 = torch.randn((3, 2, 16, 12), device=_dev(), dtype=torch.float64)
    assert torch.equal(x1, smp.sample(labels=labels, latents=lat1))
    with pytest.raises(RuntimeError):
        dp.UnconditionalSampler(net, torch.device("cpu"), (16, 12), 2, 3).sample(labels=labels)


def test_heat_empty_mask_golden():
    import dynamical_pde_diffusion_b200 as dp

    gold, e = load_golden("joint_heat.npz"), load_golden("joint_heat_emptymask.npz")
    kw = {"dx": float(gold["dx"])}
    x, losses = _ours(gold, e, 2, 1, 2, dp.heat_loss2, kw, dp.X_and_dXdt_fd, (16, 12))
    assert np.all(losses[:, 1] == 0.0)
    assert _rel(losses, e["losses"]) < XDEV and _rel(x.numpy(), e["x"]) < XDEV
    xo, lo = _oracle_on_device(gold, e, 2, 1, 2, R.heat_loss2, kw, R.X_and_dXdt_fd, (16, 12))
    assert _rel(losses, lo) < SAME and _rel(x.numpy(), xo.numpy()) < SAME


def test_llg_golden_and_same_device_oracle():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_llg.npz")
    x, losses = _ours(gold, gold, 6, 3, 4, dp.llg_loss2, {}, dp.X_and_dXdt_dummy, (16, 8))
    assert _rel(losses, gold["losses"]) < XDEV and _rel(x.numpy(), gold["x"]) < XDEV
    xo, lo = _oracle_on_device(gold, gold, 6, 3, 4, R.llg_loss2, {}, R.X_and_dXdt_dummy, (16, 8))
    assert _rel(losses, lo) < SAME and _rel(x.numpy(), xo.numpy()) < SAME


def test_llg_residual_sampler_vs_oracle():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_llg.npz")
    dx = 500e-9 / 64
    x, losses = _ours(gold, gold, 6, 3, 4, dp.llg_residual_loss, {"dx": dx}, dp.X_and_dXdt_fd, (16, 8))
    xo, lo = _oracle_on_device(gold, gold, 6, 3, 4, R.llg_residual_loss, {"dx": dx}, R.X_and_dXdt_fd, (16, 8))
    assert np.isfinite(losses).all()
    assert _rel(losses, lo) < SAME and _rel(x.numpy(), xo.numpy()) < SAME


def test_reference_function_objects_select_the_fused_path():
    """A caller passing the reference's own heat_loss2 (module ...pde_losses) gets the fused kernels."""
    import types
    from dynamical_pde_diffusion_b200.sampler import _pde_kind_of
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT, PDE_LLG_NORM

    def heat_loss2(u, dudt, labels, dx):
        raise AssertionError("must not be called")

    heat_loss2.__module__ = "diffusion_pde.sampling.pde_losses"
    assert _pde_kind_of(heat_loss2) == PDE_HEAT
    f = types.FunctionType(heat_loss2.__code__, {}, "llg_loss2")
    f.__module__ = "diffusion_pde.sampling.pde_losses"
    assert _pde_kind_of(f) == PDE_LLG_NORM
    assert _pde_kind_of(lambda *a, **k: None) is None


def test_generic_loss_fn_plugin_matches_fused():
    """An arbitrary callable in the loss_fn slot (here: the oracle's torch heat_loss2) takes the unfused route."""
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    kw = {"dx": float(gold["dx"])}
    x1, l1 = _ours(gold, gold, 2, 1, 2, dp.heat_loss2, kw, dp.X_and_dXdt_fd, (16, 12))
    x2, l2 = _ours(gold, gold, 2, 1, 2, lambda u, d, lab, dx: R.heat_loss2(u, d, lab, dx), kw, dp.X_and_dXdt_fd, (16, 12))
    assert _rel(l2, l1) < SAME and _rel(x2.numpy(), x1.numpy()) < SAME


def test_level1_plugin_inside_the_reference_loop():
    """Level-1 boundary: the reference sampler loop (oracle restatement of sample.py:320-357) with OUR heat_loss2
    plugged into its loss_fn slot, fp64 channel-slice views and all."""
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    kw = {"dx": float(gold["dx"])}
    xo, lo = _oracle_on_device(gold, gold, 2, 1, 2, R.heat_loss2, kw, R.X_and_dXdt_fd, (16, 12))
    xp, lp = _oracle_on_device(gold, gold, 2, 1, 2, dp.heat_loss2, kw, R.X_and_dXdt_fd, (16, 12))
    assert _rel(lp, lo) < SAME and _rel(xp.numpy(), xo.numpy()) < SAME


def test_jvp_time_derivative_provider_carries_dxdt_gradient():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    kw = {"dx": float(gold["dx"])}
    e = dict(gold)
    e["num_steps"] = np.int64(4)
    x, losses = _ours(gold, e, 2, 1, 2, dp.heat_loss2, kw, dp.X_and_dXdt, (16, 12))
    xo, lo = _oracle_on_device(gold, e, 2, 1, 2, R.heat_loss2, kw, R.X_and_dXdt, (16, 12))
    assert _rel(losses, lo) < SAME and _rel(x.numpy(), xo.numpy()) < SAME


def test_default_latents_are_the_first_rng_draw():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    e = dict(gold)
    e["num_steps"] = np.int64(3)
    torch.manual_seed(123)
    lat = torch.randn((3, 2, 16, 12), device=_dev(), dtype=torch.float64)       # what sample.py:314 draws
    e["latents"] = lat.cpu().numpy()
    x1, _ = _ours(gold, e, 2, 1, 2, dp.heat_loss2, {"dx": float(gold["dx"])}, dp.X_and_dXdt_fd, (16, 12))
    net = net_from_golden(gold, 2, 2, device=_dev())
    smp = dp.JointSampler(net, _dev(), (16, 12), 2, 3, 1, dp.heat_loss2, {"dx": float(gold["dx"])}, num_steps=3)
    torch.manual_seed(123)
    z = gold["zetas"]
    x2, none = smp.sample(torch.from_numpy(gold["labels"]), torch.from_numpy(gold["obs_a"]), torch.from_numpy(gold["obs_u"]),
                          torch.from_numpy(gold["mask_a"]), torch.from_numpy(gold["mask_u"]), float(z[0]), float(z[1]), float(z[2]))
    assert none is None and torch.equal(x1, x2)
    for attr in ("net", "device", "num_channels", "sample_shape", "num_samples", "ch_a", "loss_fn", "loss_kwargs",
                 "num_steps", "sigma_min", "sigma_max", "rho", "out_and_grad_fun", "dtype_f", "dtype_t"):
        assert hasattr(smp, attr), attr                                           # sample.py:261-276


def test_run_sweep_equals_direct_calls():
    """Config 4 driver (distributed.run_sweep) on one GPU: every (zeta, num_steps, chunk) item is one plain sample() call."""
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200 import distributed as D

    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2, device=_dev())
    H, W = 16, 12
    prob = dict(labels=torch.from_numpy(gold["labels"][:1]), obs_a=torch.from_numpy(gold["obs_a"]), obs_u=torch.from_numpy(gold["obs_u"]),
                mask_a=torch.from_numpy(gold["mask_a"]), mask_u=torch.from_numpy(gold["mask_u"]))
    zetas, steps = [(20.0, 0.5, 20.0), (2.0, 0.1, 5.0)], (3, 5)
    make = lambda n, N: dp.JointSampler(net, _dev(), (H, W), 2, n, 1, dp.heat_loss2, {"dx": float(gold["dx"])}, num_steps=N)
    final, n_done = D.run_sweep(make, prob, zetas, steps, total_samples=4, chunk=2, seed=3)
    assert final.shape == (2, 2, 4) and n_done == 2 * (3 + 5) * 4
    for zi, z in enumerate(zetas):
        for si, N in enumerate(steps):
            rows = []
            for s0 in (0, 2):
                lat = D.full_latents(2, 2, (H, W), 3 + 1000003 * zi + 7919 * N + s0)
                _, tr = make(2, N).sample(prob["labels"].expand(2, -1), prob["obs_a"], prob["obs_u"], prob["mask_a"], prob["mask_u"],
                                          *z, return_losses=True, latents=lat)
                rows.append(tr[-1].astype(np.float64))
            np.testing.assert_allclose(final[zi, si], np.mean(rows, axis=0), rtol=1e-6)


def test_batched_fd_provider_matches_the_unbatched_one():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2, device=_dev())
    x = torch.from_numpy(gold["net_in"]).to(_dev()).requires_grad_()
    sig, lab = torch.from_numpy(gold["net_sigma"]).to(_dev()), torch.from_numpy(gold["labels"]).to(_dev())
    x0a, da = dp.X_and_dXdt_fd(net, x, sig, lab)
    x0b, db = dp.X_and_dXdt_fd_batched(net, x, sig, lab)
    assert torch.equal(x0a, x0b) and x0b.requires_grad and not db.requires_grad
    # the quotient amplifies kernel-selection differences (batch B vs 2B) by 1/(2 eps) = 5e4: compare at that scale
    scale = float(x0a.abs().max()) * 1.2e-7 / 2e-5
    assert float((da - db).abs().max()) <= 8 * scale, (float((da - db).abs().max()), scale)


def test_sampler_refuses_cpu():
    import dynamical_pde_diffusion_b200 as dp

    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2)
    smp = dp.JointSampler(net, torch.device("cpu"), (16, 12), 2, 3, 1, dp.heat_loss2, {"dx": 0.1}, num_steps=3)
    with pytest.raises(RuntimeError):
        smp.sample(torch.from_numpy(gold["labels"]), torch.from_numpy(gold["obs_a"]), torch.from_numpy(gold["obs_u"]),
                   torch.from_numpy(gold["mask_a"]), torch.from_numpy(gold["mask_u"]), 1.0, 1.0, 1.0)
