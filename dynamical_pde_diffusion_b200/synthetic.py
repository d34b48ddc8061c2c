"""Seeded synthetic inputs of the shapes BASELINE.json names (no datasets ship with the reference).

Field statistics follow the reference's generators: heat initial conditions are sums of Gaussian
bumps (``src/diffusion_pde/pdes/heat.py:71-101``), labels are ``[t, alpha]`` with
``alpha = exp(U(-2.5, 0.5))`` (``pdes/heat.py:208,243-245``, ``datasets/dataset.py:97``), ``dx = 1/(H-1)``
(``pdes/heat.py:285``); LLG fields are unit vectors with an in-plane applied field of 0-50 mT
(``pdes/llg.py:108-109,162-164``).  Observation masks restate ``random_boundary_mask`` /
``random_interior_mask`` / ``combine_masks`` (``src/diffusion_pde/model_testing.py:12-124``): bool ``(H, W)``.
"""
from __future__ import annotations

import math

import torch


def _gen(seed):
    return seed if isinstance(seed, torch.Generator) else torch.Generator().manual_seed(int(seed))


def random_boundary_mask(H, W, *, frac_obs=0.5, n=None, generator=None, include_corners=True):
    """Random subset of the boundary ring (model_testing.py:12-57)."""
    ring = torch.zeros(H, W, dtype=torch.bool)
    ring[[0, -1], :] = True
    ring[:, [0, -1]] = True
    if not include_corners:
        ring[0, 0] = ring[0, -1] = ring[-1, 0] = ring[-1, -1] = False
    if n is None:
        n = int(frac_obs * (2 * H + 2 * W - 4))
    elif frac_obs == 1.0:
        return ring
    elif frac_obs == 0.0:
        return torch.zeros(H, W, dtype=torch.bool)
    return _subset(ring, n, generator)


def random_interior_mask(H, W, *, frac_obs=0.5, n=None, generator=None):
    """Random subset of the interior (model_testing.py:60-101)."""
    inner = torch.zeros(H, W, dtype=torch.bool)
    inner[1:-1, 1:-1] = True
    if n is None:
        n = int(frac_obs * (H - 2) * (W - 2))
    elif frac_obs == 1.0:
        return inner
    elif frac_obs == 0.0:
        return torch.zeros(H, W, dtype=torch.bool)
    return _subset(inner, n, generator)


def _subset(region, n, generator):
    cand = torch.where(region.flatten())[0]
    if n > cand.numel():
        raise ValueError(f"n={n} > candidate points={cand.numel()}")
    keep = cand[torch.randperm(cand.numel(), generator=generator)[:n]]
    out = torch.zeros_like(region)
    out.view(-1)[keep] = True
    return out


def combine_masks(*masks):
    """Logical OR (model_testing.py:104-124)."""
    if not masks:
        raise ValueError("At least one mask must be provided.")
    out = masks[0].clone()
    for m in masks[1:]:
        out |= m
    return out


def observation_masks(H, W, seed=0, interior_a=0.2, boundary_a=0.2, interior_u=0.05, boundary_u=0.05):
    g = _gen(seed)
    mask_a = combine_masks(random_interior_mask(H, W, frac_obs=interior_a, generator=g),
                           random_boundary_mask(H, W, frac_obs=boundary_a, generator=g))
    mask_u = combine_masks(random_interior_mask(H, W, frac_obs=interior_u, generator=g),
                           random_boundary_mask(H, W, frac_obs=boundary_u, generator=g))
    return mask_a, mask_u


def gaussian_bumps(n, H, W, seed=0, n_bumps=4):
    """(n, 1, H, W) fp32 smooth fields: sums of Gaussians, amplitude U(0.5,1), width U(0.03,0.15)."""
    g = _gen(seed)
    ys = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xs = torch.linspace(0, 1, W).view(1, 1, 1, W)
    amp = 0.5 + 0.5 * torch.rand(n, n_bumps, 1, 1, generator=g)
    wid = 0.03 + 0.12 * torch.rand(n, n_bumps, 1, 1, generator=g)
    cy = torch.rand(n, n_bumps, 1, 1, generator=g)
    cx = torch.rand(n, n_bumps, 1, 1, generator=g)
    f = amp * torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / (2 * wid ** 2))
    return f.sum(dim=1, keepdim=True).to(torch.float32)


def heat_problem(B, H, W, seed=0):
    """Observation (A, U), labels [t, alpha], dx and masks of one heat-equation sampling call.

    ``labels`` is one observation's label expanded over the batch, as ``test_loop`` does
    (model_testing.py:195-196).
    """
    g = _gen(seed)
    A = gaussian_bumps(1, H, W, seed=g)
    t = 0.5 * torch.rand(1, generator=g)
    alpha = torch.exp(-2.5 + 3.0 * torch.rand(1, generator=g))
    # a smoothed copy of A stands in for the solution at time t (the exact spectral solver is data
    # generation, out of scope); only its statistics matter for throughput and parity.
    U = 0.6 * A + 0.4 * torch.nn.functional.avg_pool2d(
        torch.nn.functional.pad(A, (2, 2, 2, 2), mode="replicate"), 5, stride=1)
    labels = torch.stack([t, alpha], dim=1).to(torch.float32).expand(B, -1).contiguous()
    mask_a, mask_u = observation_masks(H, W, seed=g)
    return dict(obs_a=A, obs_u=U, labels=labels, dx=1.0 / (H - 1), mask_a=mask_a, mask_u=mask_u,
                zeta_a=20.0, zeta_u=0.5, zeta_pde=20.0)


def llg_problem(B, H, W, seed=0):
    """Joint LLG call: a = initial magnetisation (3 ch), u = magnetisation at time t (3 ch)."""
    g = _gen(seed)
    def unit(seed_field):
        v = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(seed_field, (2, 2, 2, 2), mode="replicate"), 5, stride=1)
        return v / v.norm(dim=1, keepdim=True).clamp_min(1e-6)
    A = unit(torch.randn(1, 3, H, W, generator=g))
    U = unit(A + 0.2 * torch.randn(1, 3, H, W, generator=g))
    ang = 2 * math.pi * torch.rand(1, generator=g)
    mag = 50.0 * torch.rand(1, generator=g)
    field = torch.stack([mag * torch.cos(ang), mag * torch.sin(ang), torch.zeros(1)], dim=1)
    t = torch.rand(1, 1, generator=g)
    labels = torch.cat([t, field], dim=1).to(torch.float32).expand(B, -1).contiguous()
    mask_a, mask_u = observation_masks(H, W, seed=g)
    return dict(obs_a=A.to(torch.float32), obs_u=U.to(torch.float32), labels=labels, dx=500e-9 / 64,
                mask_a=mask_a, mask_u=mask_u, zeta_a=10.0, zeta_u=0.5, zeta_pde=10.0)
