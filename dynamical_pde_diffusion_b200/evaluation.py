"""Evaluation caller of the guided sampler: the reference's ``test_loop`` (``src/diffusion_pde/model_testing.py:161-239``)
and ``get_masks_from_config`` (``model_testing.py:126-158``) without Hydra / wandb, with the metrics kept on the GPU.

The reference pulls every batch of samples to the host (``sample.py:360``) and computes the error maps there, one
observation at a time (1000 observations = 45 minutes in ``nohup.out``).  Here ``sampler.sample(..., to_cpu=False)``
leaves the samples on the device, the four error maps of ``model_testing.py:208-215`` are written straight into
preallocated device tensors, and ONE device-to-host copy happens after the last observation.  With a process group
the observations are dealt round-robin to the ranks (every rank runs the full ``num_samples`` batch of its
observations -- no collective inside the loop) and the maps are all-gathered once at the end.

Metric definitions (bit-for-bit the reference's torch expressions, evaluated on the device):
    MAE[i]         = |obs - samples|.mean(dim=0)                  (C, H, W)   model_testing.py:208
    denom_abs[i]   = |obs|                                         (C, H, W)   model_testing.py:209
    denom_range[i] = obs.amax(H,W) - obs.amin(H,W)                 (C,)        model_testing.py:210
    std[i]         = samples.std(dim=0)                            (C, H, W)   model_testing.py:211
"""
from __future__ import annotations

import numpy as np
import torch

from .synthetic import combine_masks, random_boundary_mask, random_interior_mask

__all__ = ["get_masks", "observation_metrics", "test_loop", "summarize"]


def get_masks(sample_shape, interior_a, boundary_a, interior_u, boundary_u, same_interior=False, same_boundary=False,
              generator=None):
    """``get_masks_from_config`` (``model_testing.py:126-158``) with the config fields as arguments: bool (H, W) masks,
    drawn in the reference's order (interior a, boundary a, interior u, boundary u) so a seeded generator reproduces it."""
    H, W = sample_shape
    ia = random_interior_mask(H, W, frac_obs=interior_a, generator=generator)
    ba = random_boundary_mask(H, W, frac_obs=boundary_a, generator=generator)
    iu = ia if same_interior else random_interior_mask(H, W, frac_obs=interior_u, generator=generator)
    bu = ba if same_boundary else random_boundary_mask(H, W, frac_obs=boundary_u, generator=generator)
    return combine_masks(ia, ba), combine_masks(iu, bu)


def observation_metrics(obs: torch.Tensor, samples: torch.Tensor):
    """The four error maps of one observation (``model_testing.py:206-211``): obs (1, C, H, W), samples (B, C, H, W),
    both on the same device.  Returns (mae, d_abs, d_range, std)."""
    mae = (obs - samples).abs().mean(dim=0)
    d_abs = obs.abs()
    o = obs.squeeze(0)
    d_range = o.amax(dim=(-2, -1)) - o.amin(dim=(-2, -1))
    return mae, d_abs, d_range, samples.std(dim=0)


def test_loop(sampler, testloader, zeta_a, zeta_u, zeta_pde, mask_a=None, mask_u=None, max_num_samples=1000, *, group=None,
              log=None, save_path=None, seed=None, tf32=True, keep_on_device=False):
    """``model_testing.test_loop`` (``model_testing.py:162-239``): sampling runs inside :class:`sampling_context` as in
    the reference (``model_testing.py:186``) -- the net is put in ``eval()`` mode and moved to ``sampler.device``, cuDNN
    convolutions run in TF32 unless ``tf32=False``, and the net returns to the CPU afterwards unless
    ``keep_on_device=True``.  See :func:`_test_loop` for the arguments and the returned arrays."""
    from .sampler import sampling_context

    if torch.device(sampler.device).type != "cuda":
        raise RuntimeError(f"dpde_b200.evaluation.test_loop runs on CUDA devices only (got {sampler.device}); there is no CPU path")
    with sampling_context(sampler, tf32=tf32, keep_on_device=keep_on_device):
        return _test_loop(sampler, testloader, zeta_a, zeta_u, zeta_pde, mask_a, mask_u, max_num_samples, group=group, log=log,
                          save_path=save_path, seed=seed)


def _test_loop(sampler, testloader, zeta_a, zeta_u, zeta_pde, mask_a=None, mask_u=None, max_num_samples=1000, *, group=None,
               log=None, save_path=None, seed=None):
    """Evaluate ``sampler`` on the observations of ``testloader`` (any iterable of dicts with ``A`` (1,c,H,W), ``U``
    (1,c,H,W) and ``labels`` (1,label_dim) or None, as the reference's DataLoader yields with batch_size 1).

    Returns a dict of numpy arrays ``MAE``, ``denom_abs``, ``std`` (n, C, H, W) and ``denom_range`` (n, C) -- the
    contents of the reference's ``validation_data.npz`` (``model_testing.py:228-229``; written when ``save_path`` is
    given, by rank 0).  ``log(dict)`` receives the two per-observation scalars the reference sends to wandb
    (``model_testing.py:217-220``); they are read back from the device only when a logger is supplied.  With ``seed`` the
    latents of observation i come from ``Generator(seed + i)`` -- the result then does not depend on the number of ranks;
    without it they are the sampler's own ``torch.randn`` draws, as in the reference.
    """
    import torch.distributed as dist

    dev = torch.device(sampler.device)
    if dev.type != "cuda":
        raise RuntimeError(f"dpde_b200.evaluation.test_loop runs on CUDA devices only (got {dev}); there is no CPU path")
    C_, (H, W) = sampler.num_channels, sampler.sample_shape
    if mask_a is None:   # model_testing.py:174-177
        mask_a = torch.zeros(C_ // 2, H, W, dtype=torch.bool)
    if mask_u is None:
        mask_u = torch.zeros(C_ // 2, H, W, dtype=torch.bool)
    world = dist.get_world_size(group) if (group is not None or (dist.is_available() and dist.is_initialized())) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    try:
        total = min(len(testloader.dataset) if hasattr(testloader, "dataset") else len(testloader), max_num_samples)
    except TypeError:
        total = max_num_samples
    mine = [i for i in range(total) if i % world == rank]
    n_loc = len(mine)
    MAE = torch.zeros((n_loc, C_, H, W), device=dev)
    d_abs = torch.zeros((n_loc, C_, H, W), device=dev)
    d_rng = torch.zeros((n_loc, C_), device=dev)
    std = torch.zeros((n_loc, C_, H, W), device=dev)
    mask_a_d, mask_u_d = mask_a.to(dev), mask_u.to(dev)
    k = 0
    for i, batch in enumerate(testloader):
        if i >= total:
            break
        if i % world != rank:
            continue
        A, U, labels = batch["A"], batch["U"], batch.get("labels")
        if labels is not None:
            labels = labels.expand(sampler.num_samples, -1)            # model_testing.py:195-196
        A_d, U_d = A.to(dev, non_blocking=True), U.to(dev, non_blocking=True)
        extra = {}
        if seed is not None:
            from .distributed import full_latents
            extra["latents"] = full_latents(sampler.num_samples, C_, (H, W), seed + i)
        samples, _ = sampler.sample(labels=labels, obs_a=A_d, obs_u=U_d, mask_a=mask_a_d, mask_u=mask_u_d, zeta_a=zeta_a,
                                    zeta_u=zeta_u, zeta_pde=zeta_pde, return_losses=False, to_cpu=False, **extra)
        obs = torch.cat([A_d, U_d], dim=1).to(samples.dtype)
        MAE[k], d_abs[k], d_rng[k], std[k] = observation_metrics(obs, samples)
        if log is not None:
            r = d_rng[k][:, None, None]
            log({"rel MAE": float((MAE[k] / r).mean()), "sample rel std": float((std[k] / r).mean())})
        k += 1
    out = {"MAE": MAE[:k], "denom_abs": d_abs[:k], "denom_range": d_rng[:k], "std": std[:k]}
    if world > 1:   # one all-gather at the end; observation i lives on rank i % world at position i // world
        per = (total + world - 1) // world
        gathered = {}
        for name, t in out.items():
            pad = torch.zeros((per, *t.shape[1:]), device=dev)
            pad[: t.shape[0]] = t
            allr = torch.empty((world * per, *t.shape[1:]), device=dev)
            dist.all_gather_into_tensor(allr, pad.contiguous(), group=group)
            allr = allr.view(world, per, *t.shape[1:])
            gathered[name] = torch.stack([allr[i % world, i // world] for i in range(total)]) if total else allr[:0, 0]
        out = gathered
    res = {name: t.cpu().numpy() for name, t in out.items()}
    if save_path is not None and rank == 0:
        np.savez(save_path, **res)
    return res


def summarize(res):
    """Per-channel mean relative error, as the reference logs it (``model_testing.py:232-235``)."""
    rel = res["MAE"] / res["denom_range"][:, :, None, None]
    return rel.mean(axis=(0, 2, 3))
