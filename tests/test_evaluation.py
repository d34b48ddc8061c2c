"""Evaluation caller: mask generation and the error maps of model_testing.py:126-239.

CPU tests pin the host logic against the reference's mask generators (imported from /root/reference when it exists --
the functions are pure torch; their module is not importable as a whole because it pulls matplotlib / wandb, so the
three functions are extracted from the source text) and against the metric expressions restated in plain torch.
The GPU test runs test_loop on cuda:0 and repeats its arithmetic on the host from sampler.sample() outputs."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, net_from_golden

REF = "/root/reference/src/diffusion_pde/model_testing.py"


def _reference_mask_functions():
    """random_boundary_mask / random_interior_mask / combine_masks from the reference source (no module import)."""
    tree = ast.parse(open(REF).read())
    want = {"random_boundary_mask", "random_interior_mask", "combine_masks"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("shape", [(16, 12), (64, 64), (5, 9)])
def test_masks_match_the_reference_generators(shape):
    from dynamical_pde_diffusion_b200 import evaluation as E, synthetic as S

    ref = _reference_mask_functions()
    H, W = shape
    for frac in (0.0, 0.2, 0.5, 1.0):
        for name in ("random_boundary_mask", "random_interior_mask"):
            a = getattr(S, name)(H, W, frac_obs=frac, generator=torch.Generator().manual_seed(3))
            b = ref[name](H, W, frac_obs=frac, generator=torch.Generator().manual_seed(3))
            assert torch.equal(a, b), (name, frac)
    a = S.random_boundary_mask(H, W, n=3, generator=torch.Generator().manual_seed(1), include_corners=False)
    b = ref["random_boundary_mask"](H, W, n=3, generator=torch.Generator().manual_seed(1), include_corners=False)
    assert torch.equal(a, b)
    # get_masks = get_masks_from_config (model_testing.py:126-158): same draw order from one generator
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    ma, mu = E.get_masks(shape, 0.2, 0.3, 0.05, 0.1, generator=g1)
    ia = ref["random_interior_mask"](H, W, frac_obs=0.2, generator=g2)
    ba = ref["random_boundary_mask"](H, W, frac_obs=0.3, generator=g2)
    iu = ref["random_interior_mask"](H, W, frac_obs=0.05, generator=g2)
    bu = ref["random_boundary_mask"](H, W, frac_obs=0.1, generator=g2)
    assert torch.equal(ma, ref["combine_masks"](ia, ba)) and torch.equal(mu, ref["combine_masks"](iu, bu))
    ma, mu = E.get_masks(shape, 0.2, 0.3, 0.05, 0.1, same_interior=True, same_boundary=True, generator=torch.Generator().manual_seed(7))
    assert torch.equal(ma, mu)


def test_masks_edge_cases():
    from dynamical_pde_diffusion_b200 import synthetic as S

    assert S.random_interior_mask(8, 8, frac_obs=0.0).sum() == 0
    assert S.random_boundary_mask(8, 8, frac_obs=1.0).sum() == 28
    assert S.random_interior_mask(8, 8, n=36).sum() == 36
    with pytest.raises(ValueError):
        S.random_interior_mask(8, 8, n=37)
    with pytest.raises(ValueError):
        S.combine_masks()


def test_observation_metrics_are_the_reference_expressions():
    from dynamical_pde_diffusion_b200.evaluation import observation_metrics, summarize

    g = torch.Generator().manual_seed(0)
    obs, samples = torch.randn(1, 4, 6, 5, generator=g), torch.randn(7, 4, 6, 5, generator=g)
    mae, d_abs, d_range, std = observation_metrics(obs, samples)
    assert torch.equal(mae, (obs - samples).abs().mean(dim=0))                    # model_testing.py:208
    assert torch.equal(d_abs[0], obs.abs()[0])                                     # :209
    assert torch.equal(d_range, obs.squeeze(0).amax(dim=(-2, -1)) - obs.squeeze(0).amin(dim=(-2, -1)))   # :210
    assert torch.equal(std, samples.std(dim=0))                                    # :211
    res = {"MAE": mae[None].numpy(), "denom_range": d_range[None].numpy()}
    np.testing.assert_allclose(summarize(res), (mae / d_range[:, None, None]).mean(dim=(1, 2)).numpy(), rtol=1e-6)


def test_test_loop_refuses_cpu():
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200.evaluation import test_loop

    smp = dp.JointSampler(None, torch.device("cpu"), (8, 8), 2, 2, 1, dp.heat_loss2, {"dx": 0.1})
    with pytest.raises(RuntimeError, match="CUDA"):
        test_loop(smp, [], 1.0, 1.0, 1.0)


@pytest.fixture
def deterministic_fp32():
    """IEEE fp32 + deterministic cuDNN: two runs of the sampler only agree to 1e-5 with these (see test_gpu_sampler.py)."""
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


@pytest.mark.gpu
def test_test_loop_matches_host_side_metrics(deterministic_fp32):
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200 import evaluation as E

    dev = torch.device("cuda:0")
    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2, device=dev)
    H, W, B, N = 16, 12, 3, 4
    g = torch.Generator().manual_seed(5)
    loader = [{"A": torch.randn(1, 1, H, W, generator=g), "U": torch.randn(1, 1, H, W, generator=g),
               "labels": torch.tensor([[0.2 + 0.1 * i, 0.3]])} for i in range(3)]
    mask_a, mask_u = E.get_masks((H, W), 0.2, 0.2, 0.05, 0.05, generator=g)
    smp = dp.JointSampler(net, dev, (H, W), 2, B, 1, dp.heat_loss2, {"dx": 1.0 / (H - 1)}, num_steps=N)
    logged = []
    torch.manual_seed(11)
    res = E.test_loop(smp, loader, 20.0, 0.5, 20.0, mask_a=mask_a, mask_u=mask_u, max_num_samples=2, log=logged.append,
                      tf32=False, keep_on_device=True)
    assert res["MAE"].shape == (2, 2, H, W) and res["denom_range"].shape == (2, 2) and len(logged) == 2
    torch.manual_seed(11)                               # the same RNG stream -> the same latents per observation
    for i, batch in enumerate(loader[:2]):
        x, _ = smp.sample(batch["labels"].expand(B, -1), batch["A"], batch["U"], mask_a, mask_u, 20.0, 0.5, 20.0)
        obs = torch.cat([batch["A"], batch["U"]], dim=1)
        mae, d_abs, d_range, std = E.observation_metrics(obs, x)      # the reference's host-side arithmetic
        np.testing.assert_allclose(res["MAE"][i], mae.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(res["denom_abs"][i], d_abs[0].numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(res["denom_range"][i], d_range.numpy(), rtol=1e-6)
        np.testing.assert_allclose(res["std"][i], std.numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(logged[i]["rel MAE"], float((mae / d_range[:, None, None]).mean()), rtol=1e-4)
    # default masks (model_testing.py:174-177): all-zero (C/2, H, W) -> observation losses are the constant-zero branch
    res0 = E.test_loop(smp, loader, 20.0, 0.5, 20.0, max_num_samples=1, tf32=False, keep_on_device=True)
    assert np.isfinite(res0["MAE"]).all()
    # defaults follow the reference's sampling_context (sample.py:622-637): TF32 convolutions while sampling, eval() mode,
    # the net on the sampler's device during the loop and back on the CPU afterwards, precision setting restored
    net.train()
    before = torch.backends.cudnn.conv.fp32_precision
    res1 = E.test_loop(smp, loader, 20.0, 0.5, 20.0, max_num_samples=1)
    assert np.isfinite(res1["MAE"]).all() and not net.training
    assert next(net.parameters()).device.type == "cpu" and torch.backends.cudnn.conv.fp32_precision == before
