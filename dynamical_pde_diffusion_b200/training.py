"""Training-time physics loss: the reference's ``EDMHeatLoss``
(``src/diffusion_pde/models/loss.py:41-171``, the ME / SE variants of Physics-Informed Diffusion Models) with its
PDE term ``(dxdt - alpha * laplacian(x0*))**2`` evaluated and differentiated by CUDA kernels instead of an fp32
``F.conv2d`` + autograd.

``heat_residual_sq(u, dudt, alpha, dx)`` returns the per-sample sums ``sum_{c,h,w} (dudt - alpha_b lap(u))^2`` (B,);
everything around it -- noise draw, EDM weighting, the 1/(H W) factor, mean / sum, ``coeff / sigma**2`` -- is the
reference's own torch arithmetic, so shapes and broadcasting (including the reference's ``(B,) / (B,1,1,1)`` quirk at
``loss.py:146``) are reproduced exactly.  The kernels compute in fp64 from the network's fp32 output and round once;
the reference's fp32 convolution differs from that by fp32 rounding only (tests state the tolerance).
"""
from __future__ import annotations

import torch

from . import _ffi
from .ops import _DTYPES, _require_cuda, _rows_contiguous, _stream
from .sampler import X_and_dXdt_fd

__all__ = ["heat_residual_sq", "EDMHeatLoss"]


class _HeatResidualSq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, dudt, alpha, dx):
        _require_cuda(u, "u")
        if u.dtype not in (torch.float32, torch.float64):
            raise RuntimeError(f"heat_residual_sq: unsupported dtype {u.dtype}")
        if dudt is not None and (dudt.dtype != u.dtype or dudt.shape != u.shape):
            raise RuntimeError("heat_residual_sq: dudt must have the dtype and shape of u")
        B, Cu, H, W = u.shape
        uv = _rows_contiguous(u.detach())
        dv = _rows_contiguous(dudt.detach()) if dudt is not None else None
        a64 = alpha.detach().to(device=u.device, dtype=torch.float64).contiguous()
        if a64.shape != (B,):
            raise RuntimeError(f"heat_residual_sq: alpha must have shape ({B},), got {tuple(a64.shape)}")
        out = torch.empty(B, dtype=torch.float64, device=u.device)
        ws = torch.empty(max(_ffi.lib().dpde_heat_residual_sq_workspace_bytes(B, Cu, H, W), 8), dtype=torch.uint8, device=u.device)
        _ffi.call("dpde_heat_residual_sq", uv.data_ptr(), dv.data_ptr() if dv is not None else None, _DTYPES[u.dtype], B, Cu, H, W,
                  uv.stride(0), uv.stride(1), dv.stride(0) if dv is not None else 0, dv.stride(1) if dv is not None else 0,
                  a64.data_ptr(), float(dx), ws.data_ptr(), out.data_ptr(), _stream())
        ctx.save_for_backward(uv, dv if dv is not None else torch.empty(0, device=u.device), a64)
        ctx.has_d, ctx.dx = dv is not None, float(dx)
        ctx.want_d = dudt is not None and dudt.requires_grad
        return out.to(u.dtype)

    @staticmethod
    def backward(ctx, gout):
        uv, dv, a64 = ctx.saved_tensors
        B, Cu, H, W = uv.shape
        up = gout.detach().to(torch.float64).contiguous()
        g_u = torch.empty((B, Cu, H, W), dtype=uv.dtype, device=uv.device)
        g_d = torch.empty_like(g_u) if ctx.want_d else None
        _ffi.call("dpde_heat_residual_sq_vjp", uv.data_ptr(), dv.data_ptr() if ctx.has_d else None, _DTYPES[uv.dtype], B, Cu, H, W,
                  uv.stride(0), uv.stride(1), dv.stride(0) if ctx.has_d else 0, dv.stride(1) if ctx.has_d else 0, a64.data_ptr(),
                  ctx.dx, up.data_ptr(), g_u.data_ptr(), g_d.data_ptr() if g_d is not None else None, _stream())
        return g_u, g_d, None, None


def heat_residual_sq(u, dudt, alpha, dx):
    """Per-sample ``sum_{c,h,w} (dudt - alpha_b * laplacian(u, dx))**2`` -> (B,), differentiable in ``u`` and ``dudt``
    (``models/loss.py:143`` before its ``/(H W)``)."""
    if u.dim() != 4:
        raise RuntimeError(f"heat_residual_sq expects (B, C, H, W), got {tuple(u.shape)}")
    return _HeatResidualSq.apply(u, dudt, alpha, dx)


class EDMHeatLoss:
    """Drop-in for the reference's ``EDMHeatLoss`` (``models/loss.py:41-171``): same constructor, same ``__call__``
    signature and return shape.  Additive keyword: ``noise=(rnd_normal, eps)`` replaces the two ``torch.randn`` draws
    (``loss.py:129,132``) for reproducible comparisons."""

    def __init__(self, dx, pde_loss_coeff=1.0, method="joint", residual_estimation="ME", P_mean=-1.2, P_std=1.2,
                 sigma_data=0.5, reduce_method="mean", sigma_min=0.01, rho=7.0, steps=2):
        assert method in ["joint", "forward"], "method must be either 'joint' or 'forward'"
        assert residual_estimation in ["ME", "SE"], "residual_estimation must be either 'ME' or 'SE'"
        self.dx = dx
        self.pde_loss_coeff = pde_loss_coeff
        self.residual_estimation = residual_estimation
        self.P_mean = P_mean
        self.P_std = P_std
        self.sigma_data = sigma_data
        self.reduce_method = reduce_method
        self.sigma_min = sigma_min
        self.rho = rho
        self.steps = steps
        self.ch_a = 1 if method == "joint" else 0

    def two_step_sample(self, net, x, sigma_max, labels, **net_kwargs):
        """Short Euler sampler from per-sample ``sigma_max`` down to ``sigma_min`` (``loss.py:78-124``); PyTorch, it
        only calls the net."""
        B = x.shape[0]
        sigma_max = sigma_max.view(B)
        sigma_min = torch.tensor(float(self.sigma_min), device=x.device, dtype=torch.float32)
        idx = torch.arange(self.steps + 1, dtype=torch.float32, device=x.device)
        inv = 1.0 / self.rho
        sigmas = (sigma_max[:, None] ** inv + idx[None, :] / self.steps * (sigma_min ** inv - sigma_max[:, None] ** inv)) ** self.rho
        x_next = x
        for s_cur, s_next in zip(sigmas.T[:-1], sigmas.T[1:]):
            x_cur = x_next
            x_N = net(x_cur, s_cur.flatten(), labels, **net_kwargs)
            s_cur_b, s_next_b = s_cur.view(B, 1, 1, 1), s_next.view(B, 1, 1, 1)
            x_next = x_cur + (s_next_b - s_cur_b) * ((x_cur - x_N) / s_cur_b)
        return x_next

    def __call__(self, net, x, labels, run=None, global_step=None, noise=None, **kwargs):
        if x.device.type != "cuda":
            raise RuntimeError("dpde_b200.EDMHeatLoss runs on CUDA tensors only; there is no CPU path")
        if noise is None:
            rnd_normal = torch.randn([x.shape[0], 1, 1, 1], device=x.device)
            eps = torch.randn_like(x, device=x.device)
        else:
            rnd_normal, eps = (t.to(x.device) for t in noise)
        sigma = (rnd_normal * self.P_std + self.P_mean).exp()
        weight = (sigma ** 2 + self.sigma_data ** 2) / (sigma * self.sigma_data) ** 2
        n = eps * sigma
        D_yn, dxdt = X_and_dXdt_fd(net, x + n, sigma.flatten(), labels, **kwargs, no_grad=False)
        dxdt = dxdt.detach()[:, self.ch_a:, ...]
        edm_loss = weight * ((D_yn - x) ** 2)
        x_0star = D_yn if self.residual_estimation == "ME" else self.two_step_sample(net, D_yn, sigma, labels, **kwargs)
        H, W = x.shape[-2], x.shape[-1]
        u = x_0star[:, self.ch_a:, ...]
        res_sq = heat_residual_sq(u, dxdt, labels[:, 1], self.dx) / (H * W)        # sum over (C,H,W) of loss.py:143
        if self.reduce_method == "mean":
            edm_loss = edm_loss.mean(dim=(1, 2, 3))
            pde_loss = (res_sq / (u.shape[1] * H * W)) * self.pde_loss_coeff / (sigma ** 2)     # loss.py:146, broadcast as there
        elif self.reduce_method == "sum":
            edm_loss = edm_loss.sum(dim=(1, 2, 3))
            pde_loss = res_sq * self.pde_loss_coeff / (sigma ** 2)                                 # loss.py:149
        loss = edm_loss + pde_loss
        if run is not None:
            run.log({"Loss/train/batch/EDM": edm_loss.mean().item(), "Loss/train/batch/PDE": pde_loss.mean().item(),
                     "Loss/train/batch/Total": loss.mean().item()}, step=global_step)
        return loss
