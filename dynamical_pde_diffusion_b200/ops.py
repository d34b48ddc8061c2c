"""Host-side mirror of the reference's plug-in functions, backed by the sm_100a kernels.

Level 1 of the drop-in boundary: ``laplacian``, ``heat_loss2``, ``llg_loss2`` and
``llg_residual_loss`` are ``torch.autograd.Function`` s with the reference's signatures and return shapes
(``src/diffusion_pde/sampling/sample.py:106``, ``sampling/pde_losses.py:71,99``), so the UNMODIFIED reference
``JointSampler`` runs on them: forward launches the residual/reduce kernel, backward the analytic-VJP kernel.
They accept what that sampler passes: fp64 (or fp32) channel-slice views ``x_N[:, ch_a:]``, fp32 labels, numpy
``dx``, an all-zero ``dudt`` that may or may not require grad.

:class:`GuidanceEngine` is the fused form used by our :class:`~dynamical_pde_diffusion_b200.sampler.JointSampler`:
one reduce + one VJP launch produce all three losses and the seed gradient d loss_comb / d x0-hat.

Every function raises on non-CUDA tensors: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import torch

from . import _ffi
from ._ffi import F32, F64, U8, PDE_HEAT, PDE_LLG_NORM, PDE_LLG_RESIDUAL, PDE_NONE, GuidanceDesc, View

_DTYPES = {torch.float32: F32, torch.float64: F64, torch.uint8: U8}


@dataclass(frozen=True)
class LLGConstants:
    """muMAG standard problem 4 constants (``tests/test_llg_pde_loss.py:36-41``, ``pdes/llg.py:66,75-78``)."""
    gamma: float = 2.21e5
    alpha: float = 4.42e3
    A0: float = 1.3e-11
    Ms: float = 8e5
    K0: float = 0.0
    mu0: float = 4e-7 * math.pi
    t_per_step: float = 4e-12
    n_t: int = 1
    easy_axis: tuple = (1.0, 0.0, 0.0)

    @property
    def c_ex(self):
        return 2.0 * self.A0 / (self.mu0 * self.Ms)

    @property
    def c_an(self):
        return 2.0 * self.K0 / (self.mu0 * self.Ms)

    @property
    def tau(self):
        return self.t_per_step * self.n_t


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"dpde_b200: `{name}` must be a CUDA tensor (got {t.device}); these ops have no CPU path")


def _rows_contiguous(t: torch.Tensor) -> torch.Tensor:
    """Rows must be contiguous (stride_w == 1, stride_h == W); batch/channel strides are free."""
    W = t.shape[-1]
    if t.stride(-1) == 1 and t.stride(-2) == W:
        return t
    return t.contiguous()


def field_view(t: torch.Tensor) -> tuple[View, torch.Tensor]:
    """(B, ch, H, W) tensor -> dpde_view (keeps channel-slice views copy-free)."""
    t = _rows_contiguous(t)
    return View(t.data_ptr(), _DTYPES[t.dtype], 0, t.stride(0), t.stride(1)), t


def broadcast_view(t: torch.Tensor, B: int, ch: int, H: int, W: int, kind: str) -> tuple[View, torch.Tensor]:
    """Observation / mask operand broadcastable to (B, ch, H, W) -> view with 0 strides on broadcast dims.

    Accepts the shapes the reference's callers pass: (H,W) masks (``test2.py:49``), (ch,H,W) defaults
    (``model_testing.py:174-177``), (1,ch,H,W) observations (``model_testing.py:192-193``), full (B,ch,H,W).
    """
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)        # the C ABI's DPDE_U8 operands are 0 / 1 masks (the fast paths test bits, not values)
    elif t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)        # weights of any other dtype (uint8 included) keep the reference's multiply semantics
    if t.shape[-2:] != (H, W):
        t = t.expand(*t.shape[:-2], H, W).contiguous() if t.dim() >= 2 else t.expand(H, W).contiguous()
    t = _rows_contiguous(t)
    e = t.expand(B, ch, H, W)  # raises like torch broadcasting would in the reference
    return View(t.data_ptr(), _DTYPES[t.dtype], 0, e.stride(0), e.stride(1)), t


class _Scratch:
    """Per-device reduce workspace (zero-filled once) -- the only memory the ops allocate besides outputs."""
    _by_device: dict = {}

    @classmethod
    def get(cls, device: torch.device) -> torch.Tensor:
        key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
        ws = cls._by_device.get(key)
        if ws is None:
            ws = torch.zeros(_ffi.lib().dpde_guidance_workspace_bytes(), dtype=torch.uint8, device=device)
            cls._by_device[key] = ws
        return ws


def _fill_llg(desc: GuidanceDesc, consts: LLGConstants) -> None:
    desc.gamma, desc.alpha, desc.c_ex, desc.c_an, desc.tau = consts.gamma, consts.alpha, consts.c_ex, consts.c_an, consts.tau
    desc.easy_axis = (C.c_double * 3)(*consts.easy_axis)


class GuidanceEngine:
    """All three guidance losses and their seed gradient for one (B, C, H, W) problem.

    Built once per ``sample()`` call: observation / mask views, the empty-mask branches
    (``sample.py:339,341`` -- one host sync per call instead of two per step) and the per-sample coefficients are
    fixed; :meth:`seed` then costs two kernel launches and no host synchronisation.
    """

    def __init__(self, B, C_, ch_a, H, W, pde_kind, device, *, obs_a=None, mask_a=None, obs_u=None, mask_u=None,
                 sample_coef=None, dx=0.0, llg: LLGConstants | None = None, slab=None, has_flags=None):
        """``has_flags = (has_a, has_u)`` overrides the local empty-mask test (``sample.py:339,341``) with a decision
        taken elsewhere -- globally over the ranks of a coupled batch shard or of a row-slab decomposition."""
        self.B, self.C, self.ch_a, self.H, self.W, self.kind, self.device = B, C_, ch_a, H, W, pde_kind, device
        d = GuidanceDesc()
        d.B, d.C, d.ch_a, d.H, d.W, d.pde_kind = B, C_, ch_a, H, W, pde_kind
        self._keep = []
        cu = C_ - ch_a
        d.has_a = d.has_u = 0
        if has_flags is None and slab is not None:
            has_flags = (slab["has_a"], slab["has_u"])
        if mask_a is not None and ch_a > 0:
            d.has_a = int(bool((mask_a.sum() > 0).item())) if has_flags is None else int(bool(has_flags[0]))
            if d.has_a:
                d.obs_a, t1 = broadcast_view(obs_a, B, ch_a, H, W, "obs")
                d.mask_a, t2 = broadcast_view(mask_a, B, ch_a, H, W, "mask")
                self._keep += [t1, t2]
        if mask_u is not None and cu > 0:
            d.has_u = int(bool((mask_u.sum() > 0).item())) if has_flags is None else int(bool(has_flags[1]))
            if d.has_u:
                d.obs_u, t1 = broadcast_view(obs_u, B, cu, H, W, "obs")
                d.mask_u, t2 = broadcast_view(mask_u, B, cu, H, W, "mask")
                self._keep += [t1, t2]
        if sample_coef is not None:
            sample_coef = sample_coef.to(device=device, dtype=torch.float64).contiguous()
            d.sample_coef = sample_coef.data_ptr()
            self._keep.append(sample_coef)
        d.dx = float(dx)
        if pde_kind == PDE_LLG_RESIDUAL:
            _fill_llg(d, llg or LLGConstants())
        if slab is not None:
            d.slab_halo, d.slab_row0, d.slab_H_global = int(slab["halo"]), int(slab["row0"]), int(slab["H_global"])
        self.desc = d
        self.sums = torch.zeros(3, dtype=torch.float64, device=device)
        self.scalars = torch.zeros(_ffi.NUM_SCALARS, dtype=torch.float64, device=device)
        self.workspace = _Scratch.get(device)

    def _bind(self, x0, dxdt, w):
        d = self.desc
        d.x0, x0 = field_view(x0)
        if dxdt is not None:
            if dxdt.dtype != x0.dtype:
                dxdt = dxdt.to(x0.dtype)
            d.dxdt, dxdt = field_view(dxdt)
        else:
            d.dxdt = View()
        d.w_a, d.w_u, d.w_pde = float(w[0]), float(w[1]), float(w[2])
        return x0, dxdt

    def reduce(self, x0, dxdt, weights, trace_row=None, finalize=True):
        """Pass 1.  ``trace_row``: fp32 tensor slice of 4 elements receiving the step's losses (``sample.py:357``)."""
        keep = self._bind(x0, dxdt, weights)
        _ffi.call("dpde_guidance_reduce", C.byref(self.desc), self.workspace.data_ptr(), self.sums.data_ptr(), int(finalize),
                  self.scalars.data_ptr(), trace_row.data_ptr() if trace_row is not None else None, _stream())
        return keep

    def reduce_post(self, x0, dxdt, weights, mailbox):
        """Pass 1 fused with the first half of the cross-rank sum exchange: the reduce kernel's last CTA posts this
        rank's sums into every rank's mailbox over peer memory (``dpde_guidance_reduce_post``)."""
        keep = self._bind(x0, dxdt, weights)
        _ffi.call("dpde_guidance_reduce_post", C.byref(self.desc), self.workspace.data_ptr(), self.sums.data_ptr(), C.byref(mailbox), _stream())
        return keep

    def wait_finalize(self, mailbox, timeout_s, status, trace_row=None):
        """Second half: wait for every rank's post, add the slots in rank order into ``sums``, finalise the scalars."""
        _ffi.call("dpde_mailbox_wait_finalize", C.byref(self.desc), C.byref(mailbox), float(timeout_s), status.data_ptr(),
                  self.sums.data_ptr(), self.scalars.data_ptr(), trace_row.data_ptr() if trace_row is not None else None, _stream())

    def finalize(self, trace_row=None):
        _ffi.call("dpde_guidance_finalize", C.byref(self.desc), self.sums.data_ptr(), self.scalars.data_ptr(),
                  trace_row.data_ptr() if trace_row is not None else None, _stream())

    def vjp(self, x0, dxdt, weights, want_dxdt_grad=False, upstream=None):
        """Pass 2: seed gradient in the dtype of ``x0`` (contiguous (B,C,H,W)); optionally d/d dxdt too."""
        keep = self._bind(x0, dxdt, weights)
        x0c = keep[0]
        g = torch.empty((self.B, self.C, self.H, self.W), dtype=x0c.dtype, device=x0c.device)
        if self.desc.slab_H_global > 0:
            g.zero_()
        gd = torch.empty_like(g) if want_dxdt_grad else None
        if gd is not None and self.desc.slab_H_global > 0:
            gd.zero_()
        _ffi.call("dpde_guidance_vjp", C.byref(self.desc), self.scalars.data_ptr(),
                  upstream.data_ptr() if upstream is not None else None, g.data_ptr(),
                  gd.data_ptr() if gd is not None else None, _stream())
        return g, gd

    def seed(self, x0, dxdt, weights, trace_row=None, want_dxdt_grad=False, allreduce=None):
        """reduce (+ optional cross-rank all-reduce of the three sums) + vjp."""
        if allreduce is None:
            self.reduce(x0, dxdt, weights, trace_row, finalize=True)
        else:
            self.reduce(x0, dxdt, weights, None, finalize=False)
            allreduce(self.sums)
            self.finalize(trace_row)
        return self.vjp(x0, dxdt, weights, want_dxdt_grad)


# ------------------------------------------------------------------------------------------------------------
# Level-1 plug-ins (reference signatures)
# ------------------------------------------------------------------------------------------------------------
class _Laplacian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, dx):
        _require_cuda(u, "u")
        if u.dtype not in (torch.float32, torch.float64):
            raise RuntimeError(f"laplacian: unsupported dtype {u.dtype}")
        B, Cc, H, W = u.shape
        uc = u.contiguous()
        out = torch.empty_like(uc)
        _ffi.call("dpde_laplacian", uc.data_ptr(), out.data_ptr(), _DTYPES[uc.dtype], B * Cc, H, W, H * W, float(dx), 0, _stream())
        ctx.dx = float(dx)
        return out

    @staticmethod
    def backward(ctx, gout):
        g = gout.contiguous()
        B, Cc, H, W = g.shape
        out = torch.empty_like(g)
        _ffi.call("dpde_laplacian", g.data_ptr(), out.data_ptr(), _DTYPES[g.dtype], B * Cc, H, W, H * W, ctx.dx, 1, _stream())
        return out, None


def laplacian(u, dx):
    """5-point Laplacian with reflect boundaries on a ``(B, 1, H, W)`` field (``sample.py:106-134``).

    The reference's stencil weight is ``(1,1,3,3)``, so more than one channel raises there; we keep that contract.
    """
    if u.dim() != 4 or u.shape[1] != 1:
        raise RuntimeError(f"laplacian expects a (B, 1, H, W) tensor, got {tuple(u.shape)} "
                           "(the reference's conv2d weight has a single input channel)")
    return _Laplacian.apply(u, dx)


class _PdeLoss(torch.autograd.Function):
    """loss = reduce-kernel(u, dudt); backward = VJP kernel scaled by grad_output (device scalar, no sync)."""

    @staticmethod
    def forward(ctx, u, dudt, coef, kind, dx, consts):
        _require_cuda(u, "u")
        if u.dtype not in (torch.float32, torch.float64):
            raise RuntimeError(f"pde loss: unsupported dtype {u.dtype}")
        B, Cu, H, W = u.shape
        eng = GuidanceEngine(B, Cu, 0, H, W, kind, u.device, sample_coef=coef, dx=dx, llg=consts)
        ud = u.detach()
        dd = dudt.detach() if dudt is not None else None
        eng.reduce(ud, dd, (0.0, 0.0, 1.0))
        ctx.eng, ctx.has_dudt = eng, dudt is not None
        ctx.save_for_backward(ud, dd) if dd is not None else ctx.save_for_backward(ud)
        ctx.scal = eng.scalars.clone()
        return ctx.scal[2].to(u.dtype).clone()

    @staticmethod
    def backward(ctx, gout):
        saved = ctx.saved_tensors
        u, dudt = saved[0], (saved[1] if ctx.has_dudt else None)
        eng = ctx.eng
        eng.scalars.copy_(ctx.scal)
        up = gout.detach().to(torch.float64).reshape(1).contiguous()
        want_d = ctx.has_dudt and ctx.needs_input_grad[1]
        g, gd = eng.vjp(u, dudt, (0.0, 0.0, 1.0), want_dxdt_grad=want_d, upstream=up)
        return (g if ctx.needs_input_grad[0] else None), gd, None, None, None, None


def heat_loss2(u, dudt, labels, dx):
    """Heat-equation residual loss ``sqrt(sum (dudt - alpha lap(u))^2 / (H W))`` (``pde_losses.py:71-96``)."""
    alpha = labels[:, -1].detach().to(torch.float64)
    if alpha.shape[0] != u.shape[0]:
        raise RuntimeError(f"heat_loss2: labels has {alpha.shape[0]} rows for a batch of {u.shape[0]}")
    return _PdeLoss.apply(u, dudt, alpha, PDE_HEAT, float(dx), None)


def llg_loss2(m, dmdt, labels, *args):
    """Soft ``|m| = 1`` loss ``sqrt(sum (1 - |m|)^2) / (H W)`` (``pde_losses.py:99-117``); ``dmdt``/``labels`` unused."""
    if m.shape[1] != 3:
        raise RuntimeError(f"llg_loss2 expects 3 magnetisation channels, got {m.shape[1]}")
    return _PdeLoss.apply(m, None, None, PDE_LLG_NORM, 0.0, None)


def llg_residual_loss(m, dmdt, labels, dx, consts: LLGConstants = LLGConstants()):
    """LLG residual loss ``sqrt(sum r^2)/(H W)``, ``r = dmdt - tau(-gamma m x H - alpha m x (m x H))``,
    ``H = h_ext + c_ex lap(m) + c_an (m.e) e`` (``tests/test_llg_pde_loss.py:70-117`` without demag);
    the applied field in mT is ``labels[:, -3:]``."""
    if m.shape[1] != 3:
        raise RuntimeError(f"llg_residual_loss expects 3 magnetisation channels, got {m.shape[1]}")
    h_ext = labels[:, -3:].detach().to(torch.float64) / (1000 * consts.mu0)
    return _PdeLoss.apply(m, dmdt, h_ext, PDE_LLG_RESIDUAL, float(dx), consts)


heat_loss2._dpde_kind = PDE_HEAT
llg_loss2._dpde_kind = PDE_LLG_NORM
llg_residual_loss._dpde_kind = PDE_LLG_RESIDUAL
