"""Multi-GPU sampling: one process per GPU, ``torch.distributed`` for the plumbing (NCCL on the GPU box, gloo in
the CPU tests).  The reference is single-process; everything here is new.

Batch sharding (configs 2-4).  Every per-pixel operation of the step is per-sample, but each loss is ONE norm over
the whole batch (``sample.py:340-342``, ``pde_losses.py:94,116``), so there are two legitimate semantics:

* ``independent`` (default, what north_star describes): rank r samples its B/G slice exactly as a separate
  reference call with batch B/G would; no collective inside the loop, one gather of samples and loss traces at the
  end.  Results equal G reference runs of batch B/G.
* ``coupled``: the three partial sums are all-reduced (24 bytes) between the reduce and the VJP pass of every step,
  reproducing ONE reference run with batch B up to summation order.

Row slabs (config 5) live in ``slab.py``.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise the default process group from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                **({"device_id": torch.device("cuda", local)} if backend == "nccl" else {}))
    return rank, world, local


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``total`` items for ``rank`` (first ``total % world`` ranks get one more)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t, world: int, rank: int, batch: int):
    """Slice a per-sample tensor (leading dim == batch); broadcast operands (leading dim 1 / no batch dim) pass through."""
    if t is None or not torch.is_tensor(t) or t.dim() == 0 or t.shape[0] != batch:
        return t
    lo, hi = shard_bounds(batch, world, rank)
    return t[lo:hi]


def full_latents(batch, channels, shape, seed, device="cpu"):
    """One (B,C,H,W) fp64 draw from one seed, sliced per rank, so a sharded run starts where a single-GPU run would."""
    g = torch.Generator(device=device).manual_seed(int(seed))
    return torch.randn((batch, channels, *shape), generator=g, dtype=torch.float64, device=device)


def _all_gather(t: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """(world, *t.shape) stack of every rank's ``t`` (concatenated-output form: accepted by NCCL and gloo alike)."""
    t = t.contiguous()
    out = torch.empty((world * t.shape[0], *t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.view(world, *t.shape)


def gather_samples(x_local: torch.Tensor, trace_local, batch: int, group=None, device=None):
    """All-gather per-rank samples (B_r,C,H,W) and loss traces (N,4) -> ((B,C,H,W) tensor, (G,N,4) array) on every rank.

    Uneven shards are padded to the largest shard for the collective and trimmed afterwards.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return x_local, (None if trace_local is None else np.asarray(trace_local)[None])
    rank = dist.get_rank(group)
    dev = device if device is not None else x_local.device
    sizes = [shard_bounds(batch, world, r)[1] - shard_bounds(batch, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad, *x_local.shape[1:]), dtype=x_local.dtype, device=dev)
    buf[: sizes[rank]] = x_local.to(dev)
    out = _all_gather(buf, world, group)
    x = torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)
    traces = None
    if trace_local is not None:
        t = torch.as_tensor(trace_local, dtype=torch.float32).to(dev).contiguous()
        traces = _all_gather(t, world, group).cpu().numpy()
    return x.to(x_local.device), traces


def sharded_sample(sampler, labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, *, return_losses=False,
                   latents=None, seed=None, group=None, gather=True, **kw):
    """Run ``sampler.sample`` on this rank's slice of the batch and (optionally) gather everything.

    ``labels`` / per-sample observations / masks carry the GLOBAL batch; each rank takes its slice.  With
    ``sampler.coupled`` the per-step all-reduce makes the result that of one batch-B run; otherwise shards are
    independent.  Returns ``(x, traces)``: ``x`` the global (B,C,H,W) fp32 CPU tensor (local slice if
    ``gather=False``), ``traces`` a (G,N,4) array of per-rank loss traces or None.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = labels.shape[0] if labels is not None else sampler.num_samples
    if latents is None and seed is not None:
        latents = full_latents(B, sampler.num_channels, sampler.sample_shape, seed)
    sl = lambda t: shard_batch(t, world, rank, B)
    x, tr = sampler.sample(sl(labels), sl(obs_a), sl(obs_u), sl(mask_a), sl(mask_u), zeta_a, zeta_u, zeta_pde,
                           return_losses=return_losses, latents=sl(latents), **kw)
    if not gather or world == 1:
        return x, (None if tr is None else tr[None])
    dev = torch.device(sampler.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    return gather_samples(x, tr, B, group=group, device=dev)


def sweep_work_items(zeta_values, step_counts, total_samples, chunk):
    """Config 4: (zeta index, num_steps, first sample, n samples) work items; ranks take items round-robin."""
    items = []
    for zi in range(len(zeta_values)):
        for n in step_counts:
            for s0 in range(0, total_samples, chunk):
                items.append((zi, n, s0, min(chunk, total_samples - s0)))
    return items


def my_items(items, world: int, rank: int):
    return items[rank::world]


def run_sweep(make_sampler, problem, zeta_values, step_counts, total_samples, chunk, *, group=None, seed=0,
              reduce_device=None):
    """Config 4 -- the zeta / num-steps sensitivity sweep (the reference does it one call at a time in
    ``notebooks/sampler_hyperparameter_opt.ipynb``).

    ``make_sampler(num_samples, num_steps)`` returns a sampler with the ``JointSampler.sample`` signature;
    ``problem`` holds ``labels`` (1, L) (expanded to the chunk, ``model_testing.py:195-196``), ``obs_a``, ``obs_u``,
    ``mask_a``, ``mask_u``; ``zeta_values`` is a list of ``(zeta_a, zeta_u, zeta_pde)``.  The work items
    ``(zeta index, num_steps, first sample, n samples)`` are dealt round-robin to the ranks and run independently (a chunk is
    one reference call of batch ``n``, so results do not depend on the world size); the only collective is ONE
    all-reduce of the per-item summaries at the end.

    Returns ``(final_losses, n_done)``: ``final_losses[z, s]`` = the last loss-trace row ``[loss_a, loss_u, loss_pde,
    loss_comb]`` averaged over the chunks of (zeta z, step count s), shape ``(len(zeta_values), len(step_counts), 4)``,
    identical on every rank; ``n_done`` = sample-steps this rank executed (for throughput accounting).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    items = sweep_work_items(zeta_values, step_counts, total_samples, chunk)
    step_index = {n: k for k, n in enumerate(step_counts)}
    acc = torch.zeros((len(zeta_values), len(step_counts), 5), dtype=torch.float64)   # 4 loss sums + chunk count
    samplers, n_done = {}, 0
    for zi, n_steps, s0, n in my_items(items, world, rank):
        key = (n, n_steps)
        if key not in samplers:
            samplers[key] = make_sampler(n, n_steps)
        labels = problem["labels"]
        if labels is not None:
            labels = labels.expand(n, -1)
        za, zu, zp = zeta_values[zi]
        gen_seed = seed + 1000003 * zi + 7919 * n_steps + s0          # every chunk has its own reproducible latents
        lat = full_latents(n, samplers[key].num_channels, samplers[key].sample_shape, gen_seed)
        _, trace = samplers[key].sample(labels, problem["obs_a"], problem["obs_u"], problem["mask_a"], problem["mask_u"],
                                        za, zu, zp, return_losses=True, num_steps=n_steps, latents=lat)
        acc[zi, step_index[n_steps], :4] += torch.as_tensor(np.asarray(trace[-1], dtype=np.float64))
        acc[zi, step_index[n_steps], 4] += 1.0
        n_done += n * n_steps
    if world > 1:
        dev = reduce_device if reduce_device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
        t = acc.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        acc = t.cpu()
    final = (acc[..., :4] / acc[..., 4:].clamp_min(1.0)).numpy()
    return final, n_done
