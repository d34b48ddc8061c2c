"""Row-slab path on the GPU: the owned-rows update, the peer-memory halo push / flag wait kernels, and the slab sampler
run for all ranks in lock step on one device against the whole-grid sampler (same denoiser, same latents).
The true multi-process run over CUDA IPC needs >= 2 GPUs (``test_two_process_ipc``; skipped on a 1-GPU box)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dev():
    return torch.device("cuda:0")


def test_update_rows_touches_owned_rows_only_and_matches_flat():
    from dynamical_pde_diffusion_b200 import _ffi

    dev, B, C_, Hl, W, halo = _dev(), 2, 2, 14, 12, 2
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, C_, Hl, W, generator=g, dtype=torch.float64).to(dev)
    a, b, ge, gc = (torch.randn(B, C_, Hl, W, generator=g).to(dev) for _ in range(4))
    s = torch.cuda.current_stream().cuda_stream
    flat64, flat32 = torch.empty_like(x), torch.empty_like(a)
    _ffi.call("dpde_heun_guided_update", x.data_ptr(), a.data_ptr(), b.data_ptr(), ge.data_ptr(), gc.data_ptr(), 3.0, 2.0,
              flat64.data_ptr(), flat32.data_ptr(), x.numel(), s)
    for W_ in (W, W - 1):                                                # vector and scalar paths
        xs, as_, bs, ges, gcs = (t[..., :W_].contiguous() for t in (x, a, b, ge, gc))
        o64, o32 = torch.full_like(xs, 7.0), torch.full_like(as_, 7.0)
        _ffi.call("dpde_heun_guided_update_rows", xs.data_ptr(), as_.data_ptr(), bs.data_ptr(), ges.data_ptr(), gcs.data_ptr(), 3.0, 2.0,
                  o64.data_ptr(), o32.data_ptr(), B * C_, Hl * W_, halo * W_, (Hl - 2 * halo) * W_, s)
        assert torch.equal(o64[..., halo:-halo, :], flat64[..., halo:-halo, :W_])
        assert torch.equal(o32[..., halo:-halo, :], flat32[..., halo:-halo, :W_])
        for t in (o64, o32):
            assert torch.all(t[..., :halo, :] == 7.0) and torch.all(t[..., -halo:, :] == 7.0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("W", [16, 10])
def test_halo_push_and_flag_wait_single_process(dtype, W):
    """Three 'ranks' as three buffers of one device: pushes first, waits afterwards (never a wait before its push)."""
    from dynamical_pde_diffusion_b200 import _ffi
    from dynamical_pde_diffusion_b200.slab import PeerBuffer

    dev, planes, halo = _dev(), 3, 2
    Hs = [9, 8, 10]
    code = _ffi.F64 if dtype == torch.float64 else _ffi.F32
    es = 8 if dtype == torch.float64 else 4
    bufs = [PeerBuffer(planes * h * W * es + 256) for h in Hs]
    fields = [b.tensor(0, (planes, h, W), dtype, dev) for b, h in zip(bufs, Hs)]
    assert all(f.data_ptr() == b.ptr for f, b in zip(fields, bufs))
    g = torch.Generator().manual_seed(1)
    ref = [torch.randn(planes, h, W, generator=g).to(dtype).to(dev) for h in Hs]
    for f, r in zip(fields, ref):
        f.copy_(r)
    ctl = PeerBuffer(1024)                                               # flags (3 ranks x 2 sides) + ticket + status
    flags = ctl.tensor(0, (6,), torch.int64, dev)
    status = ctl.tensor(512, (1,), torch.int32, dev)
    fl = lambda rank, side: ctl.ptr + 8 * (2 * rank + side)
    s = torch.cuda.current_stream().cuda_stream
    for r in range(3):
        up, down = (r - 1 if r > 0 else None), (r + 1 if r < 2 else None)
        _ffi.call("dpde_halo_push", fields[r].data_ptr(), code, planes, Hs[r], W, halo,
                  fields[up].data_ptr() if up is not None else None, Hs[up] if up is not None else 0,
                  fields[down].data_ptr() if down is not None else None, Hs[down] if down is not None else 0,
                  fl(up, 1) if up is not None else None, fl(down, 0) if down is not None else None, 5, ctl.ptr + 256, s)
    for r in range(3):
        mine = [fl(r, side) for side, nb in ((0, r - 1), (1, r + 1)) if 0 <= nb < 3]
        arr = (C.c_void_p * len(mine))(*mine)
        _ffi.call("dpde_flag_wait", arr, len(mine), 5, 5.0, status.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert flags.tolist() == [0, 5, 5, 5, 5, 0]
    for r in range(3):
        want = ref[r].clone()
        if r > 0:
            want[:, :halo] = ref[r - 1][:, Hs[r - 1] - 2 * halo:Hs[r - 1] - halo]
        if r < 2:
            want[:, Hs[r] - halo:] = ref[r + 1][:, halo:2 * halo]
        assert torch.equal(fields[r], want), r
    # a flag nobody raises: the wait gives up after its timeout and reports it instead of hanging the stream
    arr = (C.c_void_p * 1)(fl(0, 0))
    _ffi.call("dpde_flag_wait", arr, 1, 99, 0.05, status.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(status.item()) == 1
    del fields, flags, status
    for b in bufs + [ctl]:
        b.free()


@pytest.mark.parametrize("W", [16, 10])
@pytest.mark.parametrize("last", [False, True])
def test_fused_update_and_halo_push_single_process(W, last):
    """dpde_heun_guided_update_rows_push: three 'ranks' as three buffers of one device.  Owned rows equal the flat update,
    ghost rows are untouched locally, the neighbours' ghost rows receive the freshly updated boundary rows (fp64 and
    fp32), the flags carry the epoch.  Pushes first, waits afterwards."""
    from dynamical_pde_diffusion_b200 import _ffi
    from dynamical_pde_diffusion_b200.slab import PeerBuffer

    dev, planes, halo = _dev(), 3, 2
    Hs = [9, 8, 12]
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(3)
    ctl = PeerBuffer(1024)
    flags = ctl.tensor(0, (6,), torch.int64, dev)
    status = ctl.tensor(512, (1,), torch.int32, dev)
    fl = lambda rank, side: ctl.ptr + 8 * (2 * rank + side)
    ins, outs, want = [], [], []
    for h in Hs:
        x = torch.randn(planes, h, W, generator=g, dtype=torch.float64).to(dev)
        a, b, ge, gc = (torch.randn(planes, h, W, generator=g).to(dev) for _ in range(4))
        o64, o32 = torch.full_like(x, 7.0), torch.full_like(a, 7.0)
        f64, f32 = torch.empty_like(x), torch.empty_like(a)
        _ffi.call("dpde_heun_guided_update", x.data_ptr(), a.data_ptr(), None if last else b.data_ptr(), None if last else ge.data_ptr(),
                  gc.data_ptr(), 3.0, 0.0 if last else 2.0, f64.data_ptr(), f32.data_ptr(), x.numel(), s)
        ins.append((x, a, b, ge, gc))
        outs.append((o64, o32))
        want.append((f64, f32))
    for r in range(3):
        up, down = (r - 1 if r > 0 else None), (r + 1 if r < 2 else None)
        hp = _ffi.HaloPeers()
        hp.ticket, hp.epoch = ctl.ptr + 256, 4
        if up is not None:
            hp.up64, hp.up32, hp.H_up, hp.flag_up = outs[up][0].data_ptr(), outs[up][1].data_ptr(), Hs[up], fl(up, 1)
        if down is not None:
            hp.down64, hp.down32, hp.H_down, hp.flag_down = outs[down][0].data_ptr(), outs[down][1].data_ptr(), Hs[down], fl(down, 0)
        x, a, b, ge, gc = ins[r]
        _ffi.call("dpde_heun_guided_update_rows_push", x.data_ptr(), a.data_ptr(), None if last else b.data_ptr(),
                  None if last else ge.data_ptr(), gc.data_ptr(), 3.0, 0.0 if last else 2.0, outs[r][0].data_ptr(), outs[r][1].data_ptr(),
                  planes, Hs[r], W, halo, C.byref(hp), s)
    for r in range(3):
        mine = [fl(r, side) for side, nb in ((0, r - 1), (1, r + 1)) if 0 <= nb < 3]
        arr = (C.c_void_p * len(mine))(*mine)
        _ffi.call("dpde_flag_wait", arr, len(mine), 4, 5.0, status.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(status.item()) == 0 and flags.tolist() == [0, 4, 4, 4, 4, 0]
    for r in range(3):
        for k in range(2):
            got, ref = outs[r][k], want[r][k]
            assert torch.equal(got[:, halo:Hs[r] - halo], ref[:, halo:Hs[r] - halo]), (r, k)
            top = want[r - 1][k][:, Hs[r - 1] - 2 * halo:Hs[r - 1] - halo] if r > 0 else torch.full_like(ref[:, :halo], 7.0)
            bot = want[r + 1][k][:, halo:2 * halo] if r < 2 else torch.full_like(ref[:, :halo], 7.0)
            assert torch.equal(got[:, :halo], top) and torch.equal(got[:, Hs[r] - halo:], bot), (r, k)
    del flags, status
    ctl.free()


def test_mailbox_sum_exchange_single_process():
    """dpde_guidance_reduce_post + dpde_mailbox_wait_finalize with three 'ranks' (row slabs of one grid) on one device: posts
    first, waits afterwards.  Every rank finalises the same totals, equal to the whole-grid reduce up to summation order;
    two consecutive epochs exercise both slot parities; a missing post times out and is reported."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, _ffi
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT
    from dynamical_pde_diffusion_b200.slab import PeerBuffer, SlabPlan

    dev, B, H, W, world = _dev(), 2, 36, 20, 3
    g = torch.Generator().manual_seed(5)
    x0, dxdt = torch.randn(B, 2, H, W, generator=g), 0.3 * torch.randn(B, 2, H, W, generator=g)
    obs_a, obs_u = torch.randn(1, 1, H, W, generator=g), torch.randn(1, 1, H, W, generator=g)
    mask_a, mask_u = torch.rand(H, W, generator=g) < 0.3, torch.rand(H, W, generator=g) < 0.2
    coef, dx, w = torch.rand(B, generator=g).double(), 1.0 / (H - 1), (20.0, 0.5, 20.0)
    whole = GuidanceEngine(B, 2, 1, H, W, PDE_HEAT, dev, obs_a=obs_a.to(dev), mask_a=mask_a.to(dev), obs_u=obs_u.to(dev), mask_u=mask_u.to(dev),
                           sample_coef=coef.to(dev), dx=dx)
    whole.reduce(x0.to(dev), dxdt.to(dev), w)
    boxes = [PeerBuffer(_ffi.MAILBOX_BYTES) for _ in range(world)]
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    plans = [SlabPlan(H, world, r) for r in range(world)]
    engs = [GuidanceEngine(B, 2, 1, p.H_local, W, PDE_HEAT, dev, obs_a=p.take(obs_a).to(dev), mask_a=p.take(mask_a).to(dev),
                           obs_u=p.take(obs_u).to(dev), mask_u=p.take(mask_u).to(dev), sample_coef=coef.to(dev), dx=dx,
                           slab=dict(halo=p.halo, row0=p.r0, H_global=H, has_a=True, has_u=True)) for p in plans]

    def mbox(rank, epoch):
        mb = _ffi.Mailbox()
        mb.world, mb.rank, mb.epoch = world, rank, epoch
        for r in range(world):
            mb.boxes[r] = boxes[r].ptr
        return mb

    for epoch in (1, 2):
        scale = float(epoch)                                          # different data per epoch
        for r, (p, e) in enumerate(zip(plans, engs)):
            e.reduce_post(p.take(scale * x0).contiguous().to(dev), p.take(dxdt).contiguous().to(dev), w, mbox(r, epoch))
        traces = torch.zeros(world, 4, device=dev)
        for r, e in enumerate(engs):
            e.wait_finalize(mbox(r, epoch), 5.0, status, traces[r])
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        whole.reduce(scale * x0.to(dev), dxdt.to(dev), w)
        for r, e in enumerate(engs):
            assert torch.equal(e.sums, engs[0].sums) and torch.equal(e.scalars, engs[0].scalars)      # same order on every rank
            torch.testing.assert_close(e.sums, whole.sums, rtol=1e-13, atol=0)
            torch.testing.assert_close(e.scalars[:7], whole.scalars[:7], rtol=1e-13, atol=0)
            torch.testing.assert_close(traces[r], whole.scalars[:4].float(), rtol=1e-6, atol=0)
    # epoch 3: rank 2 never posts -> the others report a timeout instead of hanging
    for r in (0, 1):
        engs[r].reduce_post(plans[r].take(x0).contiguous().to(dev), plans[r].take(dxdt).contiguous().to(dev), w, mbox(r, 3))
    engs[0].wait_finalize(mbox(0, 3), 0.05, status, None)
    torch.cuda.synchronize()
    assert int(status.item()) == 1
    for b in boxes:
        b.free()


def _heat_inputs(B, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    labels = torch.stack([0.5 * torch.rand(B, generator=g), torch.exp(-2.5 + 3 * torch.rand(B, generator=g))], 1).float()
    obs_a, obs_u = torch.randn(1, 1, H, W, generator=g), torch.randn(1, 1, H, W, generator=g)
    mask_a, mask_u = torch.rand(H, W, generator=g) < 0.3, torch.rand(H, W, generator=g) < 0.1
    lat = torch.randn(B, 2, H, W, generator=g, dtype=torch.float64)
    return labels, obs_a, obs_u, mask_a, mask_u, lat


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("provider", ["fd", "dummy"])
def test_lockstep_slab_sampler_equals_whole_grid(world, provider):
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200.slab import LockstepRanks, PointwiseDenoiser, SlabJointSampler

    dev, B, H, W, N = _dev(), 2, 40, 24, 6
    labels, obs_a, obs_u, mask_a, mask_u, lat = _heat_inputs(B, H, W)
    dx = 1.0 / (H - 1)
    net = PointwiseDenoiser().to(dev)
    fn = dp.X_and_dXdt_fd if provider == "fd" else dp.X_and_dXdt_dummy
    z = (20.0, 0.5, 20.0)
    whole = dp.JointSampler(net, dev, (H, W), 2, B, 1, dp.heat_loss2, {"dx": dx}, num_steps=N, out_and_grad_fn=fn)
    x_ref, tr_ref = whole.sample(labels, obs_a, obs_u, mask_a, mask_u, *z, return_losses=True, latents=lat)

    def make(plan, transport, allreduce):
        return SlabJointSampler(net, dev, (H, W), 2, B, 1, dp.heat_loss2, {"dx": dx}, num_steps=N, out_and_grad_fn=fn,
                                plan=plan, transport=transport, allreduce=allreduce)

    x, traces = LockstepRanks(make, H, world).sample(labels, obs_a, obs_u, mask_a, mask_u, *z, return_losses=True, latents=lat)
    assert x.shape == x_ref.shape
    # sums are added in a different order (per slab, then across slabs): agreement to fp64 round-off, not bit-exact
    assert float((x - x_ref).abs().max() / x_ref.abs().max()) < 1e-6
    for tr in traces:
        np.testing.assert_allclose(tr, tr_ref, rtol=1e-6)


def test_slab_sampler_rejects_arbitrary_loss_fn_and_cpu():
    import dynamical_pde_diffusion_b200 as dp
    from dynamical_pde_diffusion_b200.slab import PointwiseDenoiser, SlabJointSampler, SlabPlan

    labels, obs_a, obs_u, mask_a, mask_u, lat = _heat_inputs(1, 16, 8)
    s = SlabJointSampler(PointwiseDenoiser(), _dev(), (16, 8), 2, 1, 1, lambda *a, **k: 0, {}, plan=SlabPlan(16, 1, 0), transport="none")
    with pytest.raises(RuntimeError):
        s.sample(labels, obs_a, obs_u, mask_a, mask_u, 1.0, 1.0, 1.0)
    s = SlabJointSampler(PointwiseDenoiser(), "cpu", (16, 8), 2, 1, 1, dp.heat_loss2, {"dx": 0.1}, plan=SlabPlan(16, 1, 0))
    with pytest.raises(RuntimeError):
        s.sample(labels, obs_a, obs_u, mask_a, mask_u, 1.0, 1.0, 1.0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "dist"])
def test_two_process_ipc(transport):
    """Two ranks on two GPUs (torchrun, NCCL): CUDA-IPC peer pushes over NVLink vs the whole-grid run on rank 0."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "scripts", "slab_check.py"), "--transport", transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "slab_check ok" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_process_batch_shards():
    """Two ranks on two GPUs: independent shards equal per-shard calls, coupled shards equal one whole-batch call."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29651", os.path.join(ROOT, "scripts", "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "shard_check ok" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_process_two_devices():
    """The > 48 KB shared-memory opt-in of the fast-path kernels is a per-device attribute: the same process must be able
    to run them on a second GPU (heat marching kernels and LLG tiles), with the same results."""
    from dynamical_pde_diffusion_b200 import GuidanceEngine, LLGConstants
    from dynamical_pde_diffusion_b200._ffi import PDE_HEAT, PDE_LLG_RESIDUAL

    g = torch.Generator().manual_seed(2)
    B, H, W = 2, 40, 260
    outs = {}
    for kind, C_, ch_a in ((PDE_HEAT, 2, 1), (PDE_LLG_RESIDUAL, 6, 3)):
        x0, dxdt = torch.randn(B, C_, H, W, generator=g), 0.1 * torch.randn(B, C_, H, W, generator=g)
        obs_a, obs_u = torch.randn(1, ch_a, H, W, generator=g), torch.randn(1, C_ - ch_a, H, W, generator=g)
        mask = torch.rand(H, W, generator=g) < 0.2
        coef = torch.rand(B, generator=g).double() if kind == PDE_HEAT else (1e4 * torch.randn(B, 3, generator=g)).double()
        for d in (0, 1):
            dev = torch.device("cuda", d)
            with torch.cuda.device(dev):
                eng = GuidanceEngine(B, C_, ch_a, H, W, kind, dev, obs_a=obs_a.to(dev), mask_a=mask.to(dev), obs_u=obs_u.to(dev),
                                     mask_u=mask.to(dev), sample_coef=coef.to(dev), dx=1.0 / (H - 1) if kind == PDE_HEAT else 500e-9 / 64,
                                     llg=LLGConstants())
                gx, _ = eng.seed(x0.to(dev), dxdt.to(dev), (5.0, 0.5, 7.0))
                torch.cuda.synchronize(dev)
                outs[(kind, d)] = (gx.cpu(), eng.scalars[:4].cpu())
        assert torch.equal(outs[(kind, 0)][0], outs[(kind, 1)][0]) and torch.equal(outs[(kind, 0)][1], outs[(kind, 1)][1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_process_evaluation_loop():
    """evaluation.test_loop on two ranks: round-robin observations + one all-gather equal the single-process loop."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29661", os.path.join(ROOT, "scripts", "eval_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "eval_check ok" in r.stdout
