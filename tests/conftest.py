"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else runs on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def net_from_golden(gold, img_channels, label_dim, device="cpu"):
    """Rebuild the tiny denoiser stored in a fixture with OUR module tree (state-dict compatible)."""
    from dynamical_pde_diffusion_b200.denoiser import EDMPrecond, EDMUNet

    net = EDMPrecond(EDMUNet(img_channels=img_channels, label_dim=label_dim, base_channels=8, channel_mults=(1, 2),
                             num_res_blocks=1, sigma_emb_dim=8, emb_dim=16), sigma_data=0.5)
    sd = {k[len("net/"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("net/")}
    net.load_state_dict(sd, strict=True)
    return net.eval().to(device)


@pytest.fixture(scope="session")
def golden():
    return load_golden
