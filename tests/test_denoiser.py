"""Our torch denoiser against the reference network's stored output (golden) -- CPU."""
import numpy as np
import torch

from conftest import load_golden, net_from_golden


def test_denoiser_matches_reference_output():
    gold = load_golden("joint_heat.npz")
    net = net_from_golden(gold, 2, 2)
    with torch.no_grad():
        out = net(torch.from_numpy(gold["net_in"]), torch.from_numpy(gold["net_sigma"]), torch.from_numpy(gold["labels"]))
    np.testing.assert_allclose(out.numpy(), gold["net_out"], rtol=1e-5, atol=1e-6)


def test_unet_v2_parameter_count():
    from dynamical_pde_diffusion_b200.denoiser import build_unet_v2, randomize_zero_init

    net = build_unet_v2(2, 2)
    n = sum(p.numel() for p in net.parameters())
    assert abs(n - 7.04e6) < 0.02e6, n          # 7.04 M parameters for conf/model/unetv2.yaml
    zero_before = sum(int(p.abs().sum() == 0) for p in net.parameters() if p.ndim == 4)
    randomize_zero_init(net, seed=0)
    assert zero_before > 0 and all(float(p.abs().sum()) > 0 for p in net.parameters() if p.ndim == 4)
