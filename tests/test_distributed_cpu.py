"""Host-side logic of the multi-GPU paths on CPU: world_size-2 (and 3) ``gloo`` process groups.

The CUDA kernels cannot run here; what runs is everything around them -- shard bounds, batch slicing, the sample /
loss-trace gather with uneven shards, the coupled-sum all-reduce, sweep work-item assignment, the row-slab plan and
the ``torch.distributed`` halo transport with its row gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynamical_pde_diffusion_b200 import distributed as D
from dynamical_pde_diffusion_b200.slab import DistHaloExchange, SlabPlan, gather_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    try:
        out[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(world, fn):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, out), nprocs=world, join=True)
    return [out[r] for r in range(world)]


# ---- batch shards ------------------------------------------------------------------------------------------
def test_shard_bounds_cover_and_balance():
    for total in (1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            b = [D.shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _gather_job(rank, world):
    B, C_, H, W, N = 5, 2, 4, 6, 3                       # 5 samples over 2 ranks: uneven shards (3 + 2)
    full = D.full_latents(B, C_, (H, W), seed=3).float()
    lo, hi = D.shard_bounds(B, world, rank)
    labels = torch.arange(B * 2, dtype=torch.float32).view(B, 2)
    assert torch.equal(D.shard_batch(labels, world, rank, B), labels[lo:hi])
    mask = torch.ones(H, W, dtype=torch.bool)
    assert D.shard_batch(mask, world, rank, B) is mask                   # broadcast operands pass through
    trace = np.full((N, 4), float(rank), np.float32)
    x, traces = D.gather_samples(full[lo:hi] * 2, trace, B)
    return bool(torch.equal(x, full * 2)), traces.shape, traces[:, 0, 0].tolist()


def test_gather_samples_uneven_shards_gloo():
    for ok, shape, ranks in _run(2, _gather_job):
        assert ok and shape == (2, 3, 4) and ranks == [0.0, 1.0]


class _FakeSampler:
    """Stands in for JointSampler on CPU: 'samples' are a deterministic function of the slice it was handed."""
    num_channels, sample_shape, num_samples, device, coupled = 2, (4, 6), 6, "cpu", False

    def sample(self, labels, obs_a, obs_u, mask_a, mask_u, za, zu, zp, return_losses=False, latents=None, **kw):
        x = latents.float() + labels[:, :1, None, None]
        return x, (np.full((2, 4), float(labels.shape[0]), np.float32) if return_losses else None)


def _sharded_job(rank, world):
    B = 6
    labels = torch.arange(B * 2, dtype=torch.float32).view(B, 2)
    obs = torch.zeros(1, 1, 4, 6)
    mask = torch.ones(4, 6, dtype=torch.bool)
    x, tr = D.sharded_sample(_FakeSampler(), labels, obs, obs, mask, mask, 1.0, 1.0, 1.0, return_losses=True, seed=11)
    expect = D.full_latents(B, 2, (4, 6), 11).float() + labels[:, :1, None, None]
    return bool(torch.equal(x, expect)), tr.shape


def test_sharded_sample_equals_single_process_gloo():
    for ok, shape in _run(3, _sharded_job):
        assert ok and shape == (3, 2, 4)


def _coupled_job(rank, world):
    sums = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 0.5], dtype=torch.float64)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)           # what JointSampler(coupled=True) does between the two passes
    return sums.tolist()


def test_coupled_sum_allreduce_gloo():
    for got in _run(2, _coupled_job):
        assert got == [3.0, 30.0, 1.0]


def test_sweep_items_partition():
    items = D.sweep_work_items(list(range(8)), (20, 50, 200), 4096, 512)
    assert len(items) == 8 * 3 * 8
    for world in (1, 2, 4, 8):
        mine = [D.my_items(items, world, r) for r in range(world)]
        assert sorted(sum(mine, [])) == sorted(items)
        assert max(map(len, mine)) - min(map(len, mine)) <= 1
    assert sum(n for *_, n in items) == 8 * 3 * 4096


class _StubSampler:
    """Stands in for JointSampler on CPU: the last trace row encodes the call so the sweep's bookkeeping can be checked."""
    num_channels, sample_shape = 2, (4, 4)

    def __init__(self, n, n_steps):
        self.n, self.n_steps = n, n_steps

    def sample(self, labels, obs_a, obs_u, mask_a, mask_u, za, zu, zp, return_losses=False, num_steps=None, latents=None):
        assert labels.shape[0] == self.n and latents.shape == (self.n, 2, 4, 4) and num_steps == self.n_steps
        tr = np.zeros((num_steps, 4), np.float32)
        tr[-1] = [za, zu, zp, float(latents.double().mean())]
        return torch.zeros(self.n, 2, 4, 4), tr


def _sweep_job(rank, world):
    zetas = [(1.0 * k, 0.5, 2.0 * k) for k in range(1, 4)]
    prob = dict(labels=torch.tensor([[0.1, 0.2]]), obs_a=None, obs_u=None, mask_a=None, mask_u=None)
    final, n_done = D.run_sweep(lambda n, s: _StubSampler(n, s), prob, zetas, (3, 5), 10, 4, seed=1)
    return final, n_done


def test_run_sweep_is_independent_of_world_size():
    ref, n1 = _sweep_job(0, 1)                         # no process group: single rank
    assert ref.shape == (3, 2, 4) and n1 == 3 * (3 + 5) * 10
    np.testing.assert_allclose(ref[:, 0, 0], [1.0, 2.0, 3.0])
    np.testing.assert_allclose(ref[:, 1, 2], [2.0, 4.0, 6.0])
    for world in (2, 3):
        outs = _run(world, _sweep_job)
        assert sum(o[1] for o in outs) == n1
        for final, _ in outs:
            np.testing.assert_allclose(final, ref, rtol=1e-12)


# ---- row slabs ---------------------------------------------------------------------------------------------
def test_slab_plan_rows_and_take():
    H, W = 11, 5
    field = torch.arange(H * W, dtype=torch.float64).view(1, 1, H, W)
    covered = []
    for world in (1, 2, 3):
        for rank in range(world):
            p = SlabPlan(H, world, rank)
            loc = p.take(field)
            assert loc.shape == (1, 1, p.H_local, W)
            assert torch.equal(p.owned(loc), field[..., p.r0:p.r1, :])
            if p.up is None:
                assert torch.all(loc[..., :p.halo, :] == 0)              # beyond the grid: zeros, never read
            else:
                assert torch.equal(loc[..., :p.halo, :], field[..., p.r0 - p.halo:p.r0, :])
            if world == 3:
                covered += list(range(p.r0, p.r1))
    assert covered == list(range(H))
    with pytest.raises(ValueError):
        SlabPlan(6, 4, 0)                                                # slabs shorter than the halo


def _halo_job(rank, world):
    H, W, B, C_ = 13, 6, 2, 2
    g = torch.Generator().manual_seed(0)
    full64 = torch.randn(B, C_, H, W, generator=g, dtype=torch.float64)
    full32 = full64.float() * 3
    p = SlabPlan(H, world, rank)
    l64, l32 = p.take(full64), p.take(full32)
    h = p.halo
    for t in (l64, l32):                                                 # forget the ghost rows, then exchange
        t[..., :h, :] = -1
        t[..., p.H_local - h:, :] = -1
    DistHaloExchange(p).exchange(l64, l32)
    ok = True
    for loc, full in ((l64, full64), (l32, full32)):
        want = p.take(full)
        if p.up is None:
            want[..., :h, :] = -1                                        # no neighbour: untouched
        if p.down is None:
            want[..., p.H_local - h:, :] = -1
        ok = ok and bool(torch.equal(loc, want))
    gathered = gather_rows(p.owned(l32).contiguous(), p)
    return ok, bool(torch.equal(gathered, full32))


@pytest.mark.parametrize("world", [2, 3])
def test_dist_halo_exchange_and_row_gather_gloo(world):
    for ok, gathered_ok in _run(world, _halo_job):
        assert ok and gathered_ok


def test_slab_sampler_rejects_unknown_transports():
    """A typo such as 'nccl' must not fall through to a run without any halo exchange."""
    from dynamical_pde_diffusion_b200.slab import SlabJointSampler

    args = (torch.nn.Identity(), "cuda", (16, 8), 2, 1, 1, None, {})
    for bad in ("nccl", "Peer", None, object()):
        with pytest.raises(ValueError):
            SlabJointSampler(*args, plan=SlabPlan(16, 2, 0), transport=bad)
    with pytest.raises(ValueError):
        SlabJointSampler(*args, plan=SlabPlan(16, 2, 0), transport="none")      # no exchange with two ranks
    assert SlabJointSampler(*args, plan=SlabPlan(16, 1, 0), transport="none").transport_kind == "none"
    assert SlabJointSampler(*args, plan=SlabPlan(16, 2, 1), transport="dist").transport_kind == "dist"
