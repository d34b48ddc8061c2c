"""Live comparison of the oracle with the unmodified reference (skipped where /root/reference is absent)."""
import numpy as np
import pytest
import torch

from oracle import guided_sampler_ref as R
from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


def test_functions_match():
    S, PL, _ = import_reference()
    g = torch.Generator().manual_seed(0)
    u = torch.randn(4, 1, 20, 13, generator=g, dtype=torch.float64)
    d = torch.randn(4, 1, 20, 13, generator=g, dtype=torch.float64)
    lab = torch.rand(4, 2, generator=g)
    assert torch.equal(R.laplacian(u, 0.05), S.laplacian(u, 0.05))
    assert torch.equal(R.heat_loss2(u, d, lab, 0.05), PL.heat_loss2(u, d, lab, 0.05))
    m = torch.randn(2, 3, 8, 8, generator=g, dtype=torch.float64)
    assert torch.equal(R.llg_loss2(m, m, None), PL.llg_loss2(m, m, None))


def test_denoiser_is_state_dict_compatible():
    _, _, M = import_reference()
    from dynamical_pde_diffusion_b200.denoiser import EDMPrecond, EDMUNet, randomize_zero_init

    torch.manual_seed(0)
    kw = dict(img_channels=2, label_dim=2, base_channels=32, channel_mults=(1, 2, 2), num_res_blocks=2, sigma_emb_dim=16, emb_dim=32)
    ref = M.EDMWrapper(M.EDMUNet(**kw)).eval()
    randomize_zero_init(ref, seed=1)
    ours = EDMPrecond(EDMUNet(**kw)).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 2, 16, 16)
    s = torch.tensor([0.5, 10.0])
    lab = torch.rand(2, 2)
    with torch.no_grad():
        torch.testing.assert_close(ours(x, s, lab), ref(x, s, lab), rtol=1e-6, atol=1e-6)
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters())


def test_full_sampler_matches_live_reference():
    S, PL, M = import_reference()
    from dynamical_pde_diffusion_b200.denoiser import randomize_zero_init

    torch.set_num_threads(1)
    torch.manual_seed(3)
    net = M.EDMWrapper(M.EDMUNet(img_channels=2, label_dim=2, base_channels=8, channel_mults=(1, 2), num_res_blocks=1,
                                 sigma_emb_dim=8, emb_dim=16)).eval()
    randomize_zero_init(net, seed=4)
    B, H, W, N = 2, 10, 14, 7
    g = torch.Generator().manual_seed(5)
    labels = torch.rand(B, 2, generator=g)
    obs_a, obs_u = torch.randn(1, 1, H, W, generator=g), torch.randn(1, 1, H, W, generator=g)
    mask_a, mask_u = torch.rand(H, W, generator=g) < 0.3, torch.rand(H, W, generator=g) < 0.2
    torch.manual_seed(9)
    lat = torch.randn(B, 2, H, W, dtype=torch.float64)
    smp = S.JointSampler(net, torch.device("cpu"), (H, W), 2, B, 1, PL.heat_loss2, {"dx": 0.1}, num_steps=N)
    torch.manual_seed(9)
    x_ref, l_ref = smp.sample(labels, obs_a, obs_u, mask_a, mask_u, 20.0, 0.5, 20.0, return_losses=True)
    x, l = R.joint_sample(net, torch.device("cpu"), (H, W), 2, 1, R.heat_loss2, {"dx": 0.1}, labels, obs_a, obs_u, mask_a, mask_u,
                          20.0, 0.5, 20.0, num_steps=N, latents=lat)
    np.testing.assert_array_equal(l, l_ref)
    assert torch.equal(x, x_ref)
