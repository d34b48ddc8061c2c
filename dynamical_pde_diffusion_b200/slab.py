"""Row-slab domain decomposition of the guided sampler for grids too large for one GPU (BASELINE config 5:
heat 4096 x 4096, batch 8, 2/4/8 GPUs).  Nothing like it exists in the reference (single process).

Rank r owns global rows ``[r0, r1)`` of every (b, c) plane and keeps ``halo = 2`` ghost rows on each side -- what the
fused heat VJP needs (residual at +-1 row, its stencil another row) and what the LLG residual needs.  Reflect
boundaries apply only at the global top and bottom (``slab_row0`` / ``slab_H_global`` of ``dpde_guidance_desc``).

Per guided step and rank (``transport="peer"``, the product path):

    dpde_flag_wait                        ghost rows of the state pushed by the neighbours have landed (usually long ago)
    denoiser(s) + dpde_euler_predict      on the whole local buffer: a pointwise / local denoiser maps valid ghost
                                          rows of the state to valid ghost rows of x0-hat -- no exchange of x0-hat
    dpde_guidance_reduce_post             partial sums of the owned rows; the kernel's last CTA stores them into EVERY
                                          rank's mailbox through peer pointers and releases the slot's flag
    dpde_mailbox_wait_finalize            acquires the `world` slots, adds them in rank order, finalises the scalars
    dpde_guidance_vjp   (owned rows)      seed gradient; torch.autograd through the denoiser
    dpde_heun_guided_update_rows_push     ONE kernel: the owned boundary rows are updated first and stored twice (local
                                          next state + the neighbours' ghost rows over NVLink), the flags are released
                                          as soon as those CTAs have fenced, the interior rows follow -- the transfer and
                                          the neighbours' wake-up overlap the bulk of the update

No NCCL call and no host synchronisation inside the loop.  ``transport="dist"`` keeps the library baseline
(``torch.distributed`` all-reduce of the three sums + send/recv of the ghost rows; gloo in the CPU tests).

Ordering (why nobody overwrites rows a neighbour is still reading).  The state ping-pongs between two buffers.  In
step s a rank reads buffer p (ghost rows included) and writes the owned rows of buffer 1-p; its push lands in the
NEIGHBOURS' ghost rows of buffer 1-p, which they last read in step s-1.  A rank's update of step s runs after its
``wait_finalize`` of step s, i.e. after EVERY rank has posted its step-s sums -- and a rank posts them from the reduce
kernel of step s, which its stream orders after all of its step s-1 kernels.  So every reader of the old contents is
done before the first peer store of step s can happen.  The sum exchange is therefore not optional decoration: it is
the synchronisation the halo push relies on (with ``dist`` the all-reduce plays the same role).  The mailbox slots are
double-buffered by step parity for the same reason: a rank reaches exchange e + 2 only after every rank has finished
reading exchange e.

The denoiser must be *local* (its output at a row depends on the input within ``halo`` rows... in fact, with one
state exchange per step, on the same row only): the reference U-Net cannot run at this size (activations exceed
HBM, and its GroupNorm is a global statistic, ``models/nets.py:172-175``), so config 5 runs the pointwise stand-in
:class:`PointwiseDenoiser`.

Two transports move the ghost rows:

* :class:`PeerHaloExchange` -- the product path: our own kernel storing through CUDA-IPC-mapped peer pointers;
* :class:`DistHaloExchange` -- ``torch.distributed`` send/recv (NCCL or gloo): the library baseline, and what the
  world-size-2 CPU tests exercise.

:class:`LockstepRanks` runs all ranks of a decomposition inside ONE process on ONE GPU in lock step (every rank's
pushes are launched before any rank waits), which is how the single-GPU test tier covers this path.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _ffi
from .denoiser import PointwiseDenoiser  # noqa: F401  (re-exported: the stand-in denoiser of config 5)
from .distributed import _all_gather, shard_bounds
from .ops import GuidanceEngine, LLGConstants, _stream
from .sampler import F32, F64, JointSampler, _pde_kind_of
from ._ffi import PDE_HEAT, PDE_LLG_RESIDUAL, PDE_NONE

HALO = 2


# ---------------------------------------------------------------------------------------------------------
# decomposition
# ---------------------------------------------------------------------------------------------------------
class SlabPlan:
    """Which global rows a rank owns, who its neighbours are, and how global operands map to local buffers."""

    def __init__(self, H: int, world: int, rank: int, halo: int = HALO):
        if world < 1 or not (0 <= rank < world):
            raise ValueError(f"bad rank {rank} of {world}")
        self.H, self.world, self.rank, self.halo = H, world, rank, halo
        self.r0, self.r1 = shard_bounds(H, world, rank)
        if H // world < halo:      # every rank's owned rows feed a neighbour's ghost rows
            raise ValueError(f"slabs of {H // world} rows are shorter than the halo ({halo}): too many ranks for H={H}")
        self.H_local = self.r1 - self.r0 + 2 * halo
        self.up = rank - 1 if rank > 0 else None           # owns the rows above (smaller row numbers)
        self.down = rank + 1 if rank < world - 1 else None

    def rows_of(self, rank: int):
        return shard_bounds(self.H, self.world, rank)

    def local_rows(self, rank: int) -> int:
        a, b = self.rows_of(rank)
        return b - a + 2 * self.halo

    def take(self, t: torch.Tensor) -> torch.Tensor:
        """(..., H, W) -> (..., H_local, W): rows [r0 - halo, r1 + halo), zero where that leaves the grid."""
        out = torch.zeros(*t.shape[:-2], self.H_local, t.shape[-1], dtype=t.dtype, device=t.device)
        lo, hi = max(self.r0 - self.halo, 0), min(self.r1 + self.halo, self.H)
        off = lo - (self.r0 - self.halo)
        out[..., off:off + hi - lo, :] = t[..., lo:hi, :]
        return out

    def owned(self, t: torch.Tensor) -> torch.Tensor:
        return t[..., self.halo:self.H_local - self.halo, :]


# ---------------------------------------------------------------------------------------------------------
# transports
# ---------------------------------------------------------------------------------------------------------
class DistHaloExchange:
    """Ghost rows by ``torch.distributed`` point-to-point (NCCL on GPUs, gloo on CPU): the library baseline."""

    def __init__(self, plan: SlabPlan, group=None):
        self.plan, self.group = plan, group

    def exchange(self, *fields: torch.Tensor) -> None:
        """Fill the ghost rows of every ``(..., H_local, W)`` field in place from the neighbours' owned rows."""
        import torch.distributed as dist

        p, h = self.plan, self.plan.halo
        ops, landing = [], []
        for f in fields:
            if p.up is not None:
                send = f[..., h:2 * h, :].contiguous()
                recv = torch.empty_like(send)
                ops += [dist.P2POp(dist.isend, send, p.up, self.group), dist.P2POp(dist.irecv, recv, p.up, self.group)]
                landing.append((f, slice(0, h), recv))
            if p.down is not None:
                send = f[..., p.H_local - 2 * h:p.H_local - h, :].contiguous()
                recv = torch.empty_like(send)
                ops += [dist.P2POp(dist.isend, send, p.down, self.group), dist.P2POp(dist.irecv, recv, p.down, self.group)]
                landing.append((f, slice(p.H_local - h, p.H_local), recv))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for f, rows, recv in landing:
            f[..., rows, :] = recv


class PeerBuffer:
    """Device memory from ``dpde_peer_alloc`` (exportable over CUDA IPC), viewed as torch tensors."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _ffi.check(_ffi.lib().dpde_peer_alloc(nbytes, C.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes

    def handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _ffi.check(_ffi.lib().dpde_peer_export(self.ptr, buf))
        return buf.raw

    def tensor(self, offset: int, shape, dtype, device) -> torch.Tensor:
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if offset + n > self.nbytes:
            raise ValueError("PeerBuffer.tensor: view exceeds the allocation")
        typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.uint8: "|u1", torch.int64: "<i8", torch.int32: "<i4"}[dtype]
        holder = type("_CudaArray", (), {})()
        holder.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                           "data": (self.ptr + offset, False), "version": 2, "strides": None}
        holder._owner = self                       # the tensor keeps `holder` alive, `holder` keeps the allocation
        return torch.as_tensor(holder, device=device)

    def free(self):
        if self.ptr:
            _ffi.check(_ffi.lib().dpde_peer_free(self.ptr))
            self.ptr = None


def _align(n, a=256):
    return (n + a - 1) // a * a


class PeerHaloExchange:
    """State buffers in exportable memory + every rank's mapping + the per-step descriptors.

    Layout of one rank's allocation: ``x64[0] | x64[1] | x32[0] | x32[1] | flags (8 x u64) | mailbox | ticket | status``.
    Flags: index 2 k + side; side 0 = written by the upper neighbour, 1 = by the lower one.  The fused update + push
    raises the k = 0 pair once per step; the stand-alone :meth:`push` (one field per launch) raises pair k per field.
    Mailbox: ``DPDE_MAILBOX_BYTES`` -- 2 step parities x 8 source ranks x {sums[3], flag}.
    """

    N_FLAGS = 8

    def __init__(self, plan: SlabPlan, planes: int, W: int, device):
        if plan.world > _ffi.MAX_RANKS:
            raise ValueError(f"the peer-memory exchange supports up to {_ffi.MAX_RANKS} ranks (one NVSwitch domain), got {plan.world}")
        self.plan, self.planes, self.W, self.device = plan, planes, W, torch.device(device)
        self.elems = lambda Hl: planes * Hl * W
        self.offsets = self._offsets(plan.H_local)
        self.buf = PeerBuffer(self.offsets["end"])
        o, shape = self.offsets, (planes, plan.H_local, W)
        self.x64 = [self.buf.tensor(o["x64_0"], shape, F64, self.device), self.buf.tensor(o["x64_1"], shape, F64, self.device)]
        self.x32 = [self.buf.tensor(o["x32_0"], shape, F32, self.device), self.buf.tensor(o["x32_1"], shape, F32, self.device)]
        self.status = self.buf.tensor(o["status"], (1,), torch.int32, self.device)
        self.peers = {plan.rank: self.buf.ptr}   # rank -> base pointer of that rank's allocation as mapped into this process
        self._opened = []
        self.epoch = 0             # halo pushes completed so far; identical on every rank
        self.sum_epoch = 0         # sum exchanges started so far; identical on every rank
        self.timeout_s = 30.0

    def _offsets(self, Hl):
        n = self.elems(Hl)
        o, cur = {}, 0
        for name, size in (("x64_0", 8 * n), ("x64_1", 8 * n), ("x32_0", 4 * n), ("x32_1", 4 * n),
                           ("flags", 8 * self.N_FLAGS), ("mailbox", _ffi.MAILBOX_BYTES), ("ticket", 8), ("status", 8)):
            o[name] = cur
            cur = _align(cur + size)
        o["end"] = cur
        return o

    # ---- wiring -------------------------------------------------------------------------------------------
    def connect_ipc(self, group=None):
        """Exchange IPC handles over ``torch.distributed`` and map EVERY other rank's allocation (the halo push needs
        the two neighbours, the sum exchange all ranks)."""
        import torch.distributed as dist

        handles = [None] * self.plan.world
        dist.all_gather_object(handles, self.buf.handle(), group=group)
        for nb in range(self.plan.world):
            if nb != self.plan.rank:
                p = C.c_void_p()
                _ffi.check(_ffi.lib().dpde_peer_open(handles[nb], C.byref(p)))
                self.peers[nb] = p.value
                self._opened.append(p.value)
        dist.barrier(group=group)

    def connect_local(self, others: dict):
        """All ranks live in this process (``LockstepRanks``): a rank's base pointer is just its pointer."""
        for nb, other in others.items():
            self.peers[nb] = other.buf.ptr

    @property
    def connected(self) -> bool:
        return len(self.peers) == self.plan.world

    def close(self):
        """Unmap the other ranks' allocations (collective protocol: every rank closes, barrier, then :meth:`free`)."""
        for p in self._opened:
            _ffi.check(_ffi.lib().dpde_peer_close(p))
        self._opened = []
        self.peers = {self.plan.rank: self.buf.ptr}

    def free(self):
        """Release this rank's exportable allocation.  Only after every other rank has closed its mapping."""
        self.x64, self.x32, self.status = [], [], None
        self.buf.free()

    # ---- per-step descriptors -----------------------------------------------------------------------------
    def mailbox(self) -> "_ffi.Mailbox":
        """``dpde_mailbox`` of the CURRENT sum exchange (``sum_epoch``)."""
        mb = _ffi.Mailbox()
        mb.world, mb.rank, mb.epoch = self.plan.world, self.plan.rank, self.sum_epoch
        for r in range(self.plan.world):
            mb.boxes[r] = self.peers[r] + self._offsets(self.plan.local_rows(r))["mailbox"]
        return mb

    def halo_peers(self, parity: int) -> "_ffi.HaloPeers":
        """``dpde_halo_peers`` for a fused update + push that writes state buffer ``parity``; starts a new push epoch."""
        p = self.plan
        self.epoch += 1
        hp = _ffi.HaloPeers()
        hp.ticket, hp.epoch = self.buf.ptr + self.offsets["ticket"], self.epoch
        if p.up is not None:     # I am the LOWER neighbour of my upper neighbour: its "written by the lower neighbour" flag
            base, off = self.peers[p.up], self._offsets(p.local_rows(p.up))
            hp.up64, hp.up32, hp.H_up = base + off[f"x64_{parity}"], base + off[f"x32_{parity}"], p.local_rows(p.up)
            hp.flag_up = base + off["flags"] + 8 * 1
        if p.down is not None:
            base, off = self.peers[p.down], self._offsets(p.local_rows(p.down))
            hp.down64, hp.down32, hp.H_down = base + off[f"x64_{parity}"], base + off[f"x32_{parity}"], p.local_rows(p.down)
            hp.flag_down = base + off["flags"] + 8 * 0
        return hp

    def push(self, parity: int):
        """Stand-alone halo push (one launch per field) of the owned boundary rows of state buffer ``parity``: the
        unfused path for slabs shorter than 4 halo rows, and the subject of the single-kernel tests."""
        p, h = self.plan, self.plan.halo
        self.epoch += 1
        me = self.offsets
        for k, (name, dtype, t) in enumerate((("x64", _ffi.F64, self.x64[parity]), ("x32", _ffi.F32, self.x32[parity]))):
            args = {}
            for side, nb in (("up", p.up), ("down", p.down)):
                if nb is None:
                    args[side] = (None, 0, None)
                    continue
                base, off = self.peers[nb], self._offsets(p.local_rows(nb))
                # my rows land in the neighbour's ghost rows on ITS opposite side: I am its lower neighbour when it
                # is my upper one, so I raise its "written by the lower neighbour" flag (side index 1), and vice versa
                flag = base + off["flags"] + 8 * (2 * k + (1 if side == "up" else 0))
                args[side] = (base + off[f"{name}_{parity}"], p.local_rows(nb), flag)
            _ffi.call("dpde_halo_push", t.data_ptr(), dtype, self.planes, p.H_local, self.W, h,
                      args["up"][0], args["up"][1], args["down"][0], args["down"][1], args["up"][2], args["down"][2],
                      self.epoch, self.buf.ptr + me["ticket"], _stream())
        self._fields_pushed = 2

    def wait(self):
        """Block the stream until the neighbours' pushes of the current epoch have landed in my ghost rows."""
        p = self.plan
        flags = []
        for k in range(getattr(self, "_fields_pushed", 1)):
            if p.up is not None:
                flags.append(self.buf.ptr + self.offsets["flags"] + 8 * (2 * k + 0))
            if p.down is not None:
                flags.append(self.buf.ptr + self.offsets["flags"] + 8 * (2 * k + 1))
        if not flags or self.epoch == 0:
            return
        arr = (C.c_void_p * len(flags))(*flags)
        _ffi.call("dpde_flag_wait", arr, len(flags), self.epoch, self.timeout_s, self.status.data_ptr(), _stream())

    def check(self):
        """Host-side check of the wait status (one sync; call at the end of a run).  The status word is cleared so a
        later run on this exchange object starts clean."""
        if int(self.status.item()) != 0:
            self.status.zero_()
            raise _ffi.DpdeError("a peer-memory wait timed out (dpde_flag_wait / dpde_mailbox_wait_finalize): another rank never "
                                 "pushed its halo rows or posted its sums; the run's result must be discarded")


# ---------------------------------------------------------------------------------------------------------
# the sampler
# ---------------------------------------------------------------------------------------------------------
class SlabJointSampler(JointSampler):
    """``JointSampler`` on one row slab.  Same constructor and ``sample()`` arguments, all of them GLOBAL (full-grid
    observations, masks, latents); ``sample()`` returns this rank's owned rows unless ``gather=True``.

    ``plan`` fixes the decomposition; ``transport`` is ``"peer"`` (NVLink peer-memory kernel, default on CUDA with
    an initialised process group), ``"dist"`` (torch.distributed send/recv) or ``"none"`` (no exchange at all: only
    legal for a single-rank plan).  Anything else raises: a silently skipped exchange would leave the ghost rows at
    their initial values for the whole run."""

    TRANSPORTS = ("peer", "dist", "none")

    def __init__(self, *args, plan: SlabPlan, transport="peer", group=None, allreduce=None, **kw):
        super().__init__(*args, **kw)
        if transport not in self.TRANSPORTS:
            raise ValueError(f"SlabJointSampler: unknown transport {transport!r}; choose one of {self.TRANSPORTS}")
        if transport == "none" and plan.world > 1:
            raise ValueError("SlabJointSampler: transport 'none' exchanges no ghost rows and is only valid for a single-rank plan")
        self.plan, self.transport_kind, self.group = plan, transport, group
        self._allreduce_fn = allreduce
        self.peer = None

    def _allreduce(self):
        if self.transport_kind == "peer" and self.plan.world > 1:
            return self._no_allreduce       # never called: _pass1 / _combine take the mailbox path (only marks "not finalised")
        if self._allreduce_fn is not None:
            return self._allreduce_fn
        if self.plan.world == 1:
            return None
        import torch.distributed as dist
        group = self.group
        return lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    @staticmethod
    def _no_allreduce(sums):
        raise RuntimeError("SlabJointSampler: the peer transport exchanges its sums through the mailboxes; connect the ranks first")

    def begin(self, labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, num_steps=None, sigma_min=None,
              sigma_max=None, rho=None, latents=None, generator=None, connect=True):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError(f"dpde_b200.SlabJointSampler runs on CUDA devices only (got {dev}); there is no CPU path")
        plan = self.plan
        num_steps = num_steps if num_steps is not None else self.num_steps
        sigma_min = sigma_min if sigma_min is not None else self.sigma_min
        sigma_max = sigma_max if sigma_max is not None else self.sigma_max
        rho = rho if rho is not None else self.rho
        H, W = self.sample_shape
        if H != plan.H:
            raise ValueError(f"plan was made for H={plan.H}, sampler has H={H}")
        C_, ch_a = self.num_channels, self.ch_a
        kind = _pde_kind_of(self.loss_fn)
        if kind is None:
            raise RuntimeError("SlabJointSampler needs one of the fused residuals (heat_loss2, llg_loss2, llg_residual_loss): "
                               "an arbitrary loss_fn cannot be evaluated on a slab")
        sigmas = self._sigmas(num_steps, sigma_min, sigma_max, rho)
        B = labels.shape[0] if labels is not None else self.num_samples
        if labels is not None:
            labels = labels.to(device=dev, dtype=F32)

        # the empty-mask branches (sample.py:339,341) are decided on the GLOBAL masks
        has_a = bool(mask_a.sum() > 0) if ch_a > 0 else False
        has_u = bool(mask_u.sum() > 0) if C_ - ch_a > 0 else False

        def local(t, ch):
            return plan.take(t.to(dev)).contiguous()      # bool masks stay bool: they reach the kernels as 0 / 1 bytes

        coef, dx, llg = None, 0.0, None
        if kind == PDE_HEAT:
            coef, dx = labels[:, -1].to(F64), float(self.loss_kwargs["dx"])
        elif kind == PDE_LLG_RESIDUAL:
            llg = self.loss_kwargs.get("consts", LLGConstants())
            coef, dx = labels[:, -3:].to(F64) / (1000 * llg.mu0), float(self.loss_kwargs["dx"])
        engine = GuidanceEngine(B, C_, ch_a, plan.H_local, W, kind, dev, obs_a=local(obs_a, ch_a), mask_a=local(mask_a, ch_a),
                                obs_u=local(obs_u, C_ - ch_a), mask_u=local(mask_u, C_ - ch_a), sample_coef=coef, dx=dx, llg=llg,
                                slab=dict(halo=plan.halo, row0=plan.r0, H_global=H, has_a=has_a, has_u=has_u))

        # state: the same global draw on every rank (one seed), cut to the local rows -- ghost rows start valid
        if latents is None:
            latents = torch.randn((B, C_, H, W), device=dev, dtype=F64, generator=generator)
        lat = plan.take(latents.to(device=dev, dtype=F64)).contiguous()
        del latents

        shape = (B, C_, plan.H_local, W)
        if self.transport_kind == "peer":
            if self.peer is None or self.peer.planes != B * C_ or self.peer.W != W:
                self.peer = PeerHaloExchange(plan, B * C_, W, dev)
                if connect and plan.world > 1:
                    self.peer.connect_ipc(self.group)
            x64 = [t.view(shape) for t in self.peer.x64]
            x32 = [t.view(shape) for t in self.peer.x32]
        else:
            x64 = [torch.zeros(shape, dtype=F64, device=dev) for _ in range(2)]
            x32 = [torch.zeros(shape, dtype=F32, device=dev) for _ in range(2)]
            if self.transport_kind == "dist":
                self._dist = DistHaloExchange(plan, self.group)
        _ffi.call("dpde_sampler_init", lat.data_ptr(), sigmas[0], x64[0].data_ptr(), x32[0].data_ptr(), lat.numel(), _stream())
        self._run = dict(engine=engine, fused=True, sigmas=sigmas, N=num_steps, B=B, labels=labels, x64=x64[0], x64_alt=x64[1],
                         x32=x32[0], x32_alt=x32[1], zetas=(zeta_a, zeta_u, zeta_pde), parity=0,
                         trace=torch.zeros((num_steps, 4), dtype=F32, device=dev), i=0, allreduce=self._allreduce())
        return self._run

    @property
    def _peer_exchange(self) -> bool:
        """The peer-memory product path: mailbox sum exchange + halo push through mapped pointers."""
        return self.transport_kind == "peer" and self.plan.world > 1 and self.peer is not None and self.peer.connected

    # ---- the three sums: reduce kernel posts into every rank's mailbox; one tiny kernel waits, adds, finalises -----
    def _pass1(self, r, engine, xN, dx_c, w, i):
        if not self._peer_exchange:
            return super()._pass1(r, engine, xN, dx_c, w, i)
        self.peer.sum_epoch += 1
        engine.reduce_post(xN, dx_c, w, self.peer.mailbox())

    def _combine(self, r, engine, i):
        if not self._peer_exchange:
            return super()._combine(r, engine, i)
        engine.wait_finalize(self.peer.mailbox(), self.peer.timeout_s, self.peer.status, r["trace"][i])

    # owned rows only: ghost rows belong to the neighbours' pushes
    def _launch_update(self, x64, x0_1c, x0_2, g_eu, g_cur, s_cur, s_next, x64n, x32n):
        p, W = self.plan, self.sample_shape[1]
        B, C_ = x64n.shape[0], x64n.shape[1]
        ptr = lambda t: t.data_ptr() if t is not None else None
        self._fused_push = self._peer_exchange and p.H_local >= 4 * p.halo
        if self._fused_push:     # update + halo push in one kernel; the next state is buffer parity ^ 1
            hp = self.peer.halo_peers(self._run["parity"] ^ 1)
            self.peer._fields_pushed = 1
            _ffi.call("dpde_heun_guided_update_rows_push", x64.data_ptr(), x0_1c.data_ptr(), ptr(x0_2), ptr(g_eu), ptr(g_cur), s_cur, s_next,
                      x64n.data_ptr(), x32n.data_ptr(), B * C_, p.H_local, W, p.halo, C.byref(hp), _stream())
            return
        _ffi.call("dpde_heun_guided_update_rows", x64.data_ptr(), x0_1c.data_ptr(), ptr(x0_2), ptr(g_eu), ptr(g_cur),
                  s_cur, s_next, x64n.data_ptr(), x32n.data_ptr(), B * C_, p.H_local * W, p.halo * W,
                  (p.H_local - 2 * p.halo) * W, _stream())

    def _step_back(self, ctx):
        super()._step_back(ctx)
        self._run["parity"] ^= 1
        self.exchange_push()

    def exchange_push(self):
        """Send the freshly written owned boundary rows of the current state to the neighbours (already done by the
        fused update + push kernel on the peer path)."""
        r = self._run
        if self.plan.world == 1:
            return
        if self.transport_kind == "peer":
            if not getattr(self, "_fused_push", False):
                self.peer.push(r["parity"])
        elif self.transport_kind == "dist":
            self._dist.exchange(r["x64"], r["x32"])

    def exchange_wait(self):
        if self.plan.world > 1 and self.transport_kind == "peer":
            self.peer.wait()

    def step(self):
        self.exchange_wait()
        ctx = self._step_front()
        self._step_back(ctx)

    def finish(self, return_losses=False, gather=False):
        r = self._run
        self.exchange_wait()
        x = self.plan.owned(r["x32"].detach()).contiguous()
        losses = r["trace"].cpu().numpy() if return_losses else None
        if self.transport_kind == "peer" and self.peer is not None and self.plan.world > 1:
            self.peer.check()
        self._run = None
        if gather and self.plan.world > 1:
            x = gather_rows(x, self.plan, self.group)
        return x.cpu(), losses

    def release(self):
        """Unmap the neighbours' buffers, then (after a barrier: every mapping must be closed before its owner frees)
        free this rank's exportable allocation."""
        if self.peer is None:
            return
        self.peer.close()
        if self.plan.world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self.group)
        self.peer.free()
        self.peer = None

    def sample(self, labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, return_losses=False,
               num_steps=None, sigma_min=None, sigma_max=None, rho=None, *, latents=None, generator=None, gather=False):
        run = self.begin(labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, num_steps, sigma_min, sigma_max,
                         rho, latents, generator)
        for _ in range(run["N"]):
            self.step()
        return self.finish(return_losses, gather=gather)


def gather_rows(x_owned: torch.Tensor, plan: SlabPlan, group=None) -> torch.Tensor:
    """All-gather the ranks' owned rows into the full ``(..., H, W)`` field (uneven slabs padded for the collective)."""
    sizes = [plan.rows_of(r)[1] - plan.rows_of(r)[0] for r in range(plan.world)]
    pad = max(sizes)
    buf = torch.zeros(*x_owned.shape[:-2], pad, x_owned.shape[-1], dtype=x_owned.dtype, device=x_owned.device)
    buf[..., : sizes[plan.rank], :] = x_owned
    out = _all_gather(buf, plan.world, group)
    return torch.cat([out[r][..., : sizes[r], :] for r in range(plan.world)], dim=-2)


# ---------------------------------------------------------------------------------------------------------
# all ranks in one process, one GPU: lock-step driver (tests, single-GPU what-if runs)
# ---------------------------------------------------------------------------------------------------------
class LockstepRanks:
    """Run the ``world`` slab samplers of one decomposition in ONE process on ONE device, phase by phase: every
    rank's front half, one sum over the ranks' partial sums, every rank's back half (which ends with its halo
    pushes), and only then the waits -- so no kernel ever waits for a launch that has not been issued."""

    def __init__(self, make_sampler, H: int, world: int, transport="peer"):
        self.world = world
        self.total = None
        self.samplers = [make_sampler(SlabPlan(H, world, r), transport, self._allreduce) for r in range(world)]

    def _allreduce(self, sums):
        sums.copy_(self.total)

    def sample(self, labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, return_losses=False, latents=None, **kw):
        runs = [s.begin(labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, latents=latents, connect=False, **kw)
                for s in self.samplers]
        peers = {r: s.peer for r, s in enumerate(self.samplers) if s.peer is not None}
        for s in self.samplers:
            if s.peer is not None:
                s.peer.connect_local(peers)
        for _ in range(runs[0]["N"]):
            ctxs = [s._step_front() for s in self.samplers]
            self.total = torch.stack([s._run["engine"].sums for s in self.samplers]).sum(0)
            for s, c in zip(self.samplers, ctxs):
                s._step_back(c)
            for s in self.samplers:
                s.exchange_wait()
        outs = [s.finish(return_losses) for s in self.samplers]
        x = torch.cat([o[0] for o in outs], dim=-2)
        return x, [o[1] for o in outs]
