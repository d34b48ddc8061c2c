"""Torch restatement of the reference's physics-guided EDM sampler (TEST INFRASTRUCTURE ONLY).

Every function restates one piece of the reference's hot path with the same ATen
operations and the same precision split (state / losses / stencil in fp64, denoiser
in fp32) so it can serve (a) as the parity oracle for the CUDA kernels and (b) as the
timed CPU baseline ("port") of ``bench.py``.  Citations are ``file:line`` relative to
``/root/reference``.  The restatement is pinned against the unmodified reference by
``tests/golden`` (written by ``oracle/make_golden.py``) and, where ``/root/reference``
exists, by a live comparison.

The product package never imports this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

F32 = torch.float32
F64 = torch.float64


# ----------------------------------------------------------------------------------------
# noise schedule                                         src/diffusion_pde/sampling/sample.py:305-308
# ----------------------------------------------------------------------------------------
def karras_sigmas(num_steps, sigma_min, sigma_max, rho, device, net=None):
    """sigma_i = (smax^(1/rho) + i/(N-1) (smin^(1/rho) - smax^(1/rho)))^rho, i<N; sigma_N = 0."""
    idx = torch.arange(num_steps, dtype=F64, device=device)
    inv = 1.0 / rho
    sig = (sigma_max ** inv + idx / (num_steps - 1) * (sigma_min ** inv - sigma_max ** inv)) ** rho
    rounder = getattr(net, "round_sigma", None)
    if rounder is not None:
        sig = rounder(sig)
    return torch.cat([sig, torch.zeros_like(sig[:1])])


# ----------------------------------------------------------------------------------------
# 5-point Laplacian with reflect padding                 src/diffusion_pde/sampling/sample.py:106-134
# ----------------------------------------------------------------------------------------
def laplacian(u, dx):
    """(u[i+1,j]+u[i-1,j]+u[i,j+1]+u[i,j-1]-4u[i,j])/dx^2 on a (B,1,H,W) field.

    Reflect padding mirrors without repeating the edge (u[-1] = u[1]); the stencil
    weight tensor is (1,1,3,3) so the channel dimension must be 1 (sample.py:126-133).
    """
    k = torch.tensor([[0, 1, 0], [1, -4, 1], [0, 1, 0]], dtype=u.dtype, device=u.device)
    padded = F.pad(u, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(padded, k[None, None]) / (dx ** 2)


def laplacian_numpy(u, dx):
    """Index-arithmetic statement of the same stencil (independent of conv2d); u: (...,H,W)."""
    u = np.asarray(u, dtype=np.float64)
    up = np.concatenate([u[..., 1:2, :], u[..., :-1, :]], axis=-2)
    dn = np.concatenate([u[..., 1:, :], u[..., -2:-1, :]], axis=-2)
    lf = np.concatenate([u[..., :, 1:2], u[..., :, :-1]], axis=-1)
    rt = np.concatenate([u[..., :, 1:], u[..., :, -2:-1]], axis=-1)
    return (up + dn + lf + rt - 4.0 * u) / (dx ** 2)


def laplacian_adjoint_numpy(g, dx):
    """Transpose of :func:`laplacian_numpy` (the stencil is not self-adjoint at the edges).

    A neighbour q contributes g[q] with weight 2 when q lies on the boundary line of that
    axis and the target is its inward neighbour, weight 1 otherwise, nothing from outside
    (SURVEY.md section 8 row a-2).
    """
    g = np.asarray(g, dtype=np.float64)
    H, W = g.shape[-2:]
    out = -4.0 * g
    wr = np.ones(H); wr[0] = 2.0; wr[-1] = 2.0
    wc = np.ones(W); wc[0] = 2.0; wc[-1] = 2.0
    gr = g * wr[:, None]
    gc = g * wc[None, :]
    out[..., 1:, :] += gr[..., :-1, :]     # from the row above
    out[..., :-1, :] += gr[..., 1:, :]     # from the row below
    out[..., :, 1:] += gc[..., :, :-1]
    out[..., :, :-1] += gc[..., :, 1:]
    return out / (dx ** 2)


# ----------------------------------------------------------------------------------------
# PDE residual losses                              src/diffusion_pde/sampling/pde_losses.py:71-117
# ----------------------------------------------------------------------------------------
def heat_loss2(u, dudt, labels, dx):
    """sqrt( sum_{b,h,w} (dudt - alpha_b lap(u))^2 / (H W) ), alpha = labels[:, -1] (pde_losses.py:91-94)."""
    alpha = labels[:, -1].view(u.shape[0], 1, 1, 1)
    r = dudt - alpha * laplacian(u, dx)
    return torch.sqrt(torch.sum(r ** 2) / (u.shape[-1] * u.shape[-2]))


def llg_loss2(m, dmdt, labels, *args):
    """Soft |m| = 1 constraint: sqrt(sum (1 - |m|)^2) / (H W) (pde_losses.py:115-116)."""
    n = torch.linalg.norm(m, dim=1)
    return torch.sqrt(torch.sum((1 - n) ** 2)) / (m.shape[2] * m.shape[3])


@dataclass(frozen=True)
class LLGConstants:
    """Material constants of muMAG standard problem 4 (tests/test_llg_pde_loss.py:36-41,
    pde_losses.py:185-191, pdes/llg.py:66,75-78)."""
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
    def c_ex(self):  # exchange prefactor 2 A0 / (mu0 Ms)      tests/test_llg_pde_loss.py:84
        return 2.0 * self.A0 / (self.mu0 * self.Ms)

    @property
    def c_an(self):  # uniaxial anisotropy prefactor 2 K0 / (mu0 Ms); reference has K0 = 0 (:87)
        return 2.0 * self.K0 / (self.mu0 * self.Ms)

    @property
    def tau(self):   # rhs scale t_per_step * n_t                tests/test_llg_pde_loss.py:117
        return self.t_per_step * self.n_t


def llg_residual_field(m, dmdt, field_mT, dx, consts: LLGConstants = LLGConstants()):
    """r = dmdt - tau * ( -gamma m x H - alpha m x (m x H) ),  H = h_ext + c_ex lap(m) + H_anis.

    Follows the torch-native "Option 1" of tests/test_llg_pde_loss.py:70-117 batched over B
    (identical algebra at pde_losses.py:194,246-250):  h_ext = field_mT / (1000 mu0) (:70),
    h_exch = c_ex * laplacian(m.unsqueeze(1), dx) per component (:82-84), h_anis = 0 for
    K0 = 0 (:87).  Demagnetisation (:90-107) needs the MagTense solver and is out of scope.
    For K0 != 0 we use the uniaxial field c_an (m . e) e  -- parity unpinned (the reference
    gets it from MagTense).
    """
    B, _, H, W = m.shape
    h_ext = field_mT.to(m.dtype).view(B, 3, 1, 1) / (1000 * consts.mu0)
    lap = laplacian(m.reshape(B * 3, 1, H, W), dx).reshape(B, 3, H, W)
    h_eff = h_ext + consts.c_ex * lap
    if consts.K0 != 0.0:
        e = torch.tensor(consts.easy_axis, dtype=m.dtype, device=m.device).view(1, 3, 1, 1)
        h_eff = h_eff + consts.c_an * (m * e).sum(dim=1, keepdim=True) * e
    mxh = torch.cross(m, h_eff, dim=1)
    rhs = -consts.gamma * mxh - consts.alpha * torch.cross(m, mxh, dim=1)
    return dmdt - rhs * consts.tau


def llg_residual_loss(m, dmdt, labels, dx, consts: LLGConstants = LLGConstants()):
    """Scalar sampler loss of the m x H_eff residual: sqrt(sum r^2) / (H W).

    The reference stops at the per-pixel field |r|_2 / n_magnets (tests/test_llg_pde_loss.py:117,
    pde_losses.py:250) and never feeds it to JointSampler; the global reduction is chosen by
    analogy with llg_loss2 (pde_losses.py:116) -- parity unpinned for this reduction only.
    The applied field is the last three label columns (heat uses labels[:, -1] the same way).
    """
    r = llg_residual_field(m, dmdt, labels[:, -3:], dx, consts)
    return torch.sqrt(torch.sum(r ** 2)) / (m.shape[2] * m.shape[3])


# ----------------------------------------------------------------------------------------
# masked observation losses                              src/diffusion_pde/sampling/sample.py:336-342
# ----------------------------------------------------------------------------------------
def obs_losses(x_N, obs_a, obs_u, mask_a, mask_u, ch_a):
    """loss = sqrt(sum (mask (x - obs))^2) per channel group; constant 0 when the mask is empty."""
    loss_u = torch.zeros(1, dtype=F64, device=x_N.device)
    loss_a = torch.zeros(1, dtype=F64, device=x_N.device)
    if mask_u.sum() > 0:
        loss_u = torch.sqrt(torch.sum((mask_u * (x_N[:, ch_a:] - obs_u)) ** 2))
    if mask_a.sum() > 0:
        loss_a = torch.sqrt(torch.sum((mask_a * (x_N[:, :ch_a] - obs_a)) ** 2))
    return loss_a, loss_u


def guidance_weights(i, num_steps, zeta_a, zeta_u, zeta_pde):
    """Observation weights drop to 10 % once i > 0.8 N, in Python floats (sample.py:348-351)."""
    if i <= 0.8 * num_steps:
        return zeta_a, zeta_u, zeta_pde
    return 0.1 * zeta_a, 0.1 * zeta_u, zeta_pde


# ----------------------------------------------------------------------------------------
# denoiser + time-derivative providers                    src/diffusion_pde/sampling/sample.py:15-103
# ----------------------------------------------------------------------------------------
def X_and_dXdt_dummy(net, x, sigma, labels, **kw):
    out = net(x, sigma, labels, **kw)
    return out, torch.zeros_like(out)


def X_and_dXdt_fd(net, x, sigma, labels, eps=1e-5, no_grad=True, **kw):
    """Central difference in labels[:, 0]; the +-eps evaluations carry no graph (sample.py:50-66)."""
    if labels is None:
        return X_and_dXdt_dummy(net, x, sigma, labels, **kw)
    lp = labels.detach().clone()
    lm = labels.detach().clone()
    lp[:, 0] += eps
    lm[:, 0] -= eps
    if no_grad:
        with torch.no_grad():
            up = net(x, sigma, lp, **kw)
            um = net(x, sigma, lm, **kw)
    else:
        up = net(x, sigma, lp, **kw)
        um = net(x, sigma, lm, **kw)
    d = (up - um) / (2 * eps)
    return net(x, sigma, labels, **kw), d


# ----------------------------------------------------------------------------------------
# one guided Heun step and the full loop               src/diffusion_pde/sampling/sample.py:320-357
# ----------------------------------------------------------------------------------------
def guided_step(net, x_in, i, sigmas, labels, obs_a, obs_u, mask_a, mask_u, ch_a,
                loss_fn, loss_kwargs, zeta_a, zeta_u, zeta_pde, num_steps,
                out_and_grad_fn=X_and_dXdt_fd, return_internals=False):
    """x_in (fp64) -> x_next (fp64), [loss_a, loss_u, loss_pde, loss_comb] (Python floats)."""
    dev = x_in.device
    B = x_in.shape[0]
    s_cur, s_next = sigmas[i], sigmas[i + 1]
    x_cur = x_in.detach().clone()
    x_cur.requires_grad = True
    x_N, dxdt = out_and_grad_fn(net, x_cur.to(F32), torch.full((B,), s_cur, device=dev, dtype=F32), labels)
    x_N, dxdt = x_N.to(F64), dxdt.to(F64)
    d_cur = (x_cur - x_N) / s_cur
    x_next = x_cur + (s_next - s_cur) * d_cur
    if i < num_steps - 1:                                        # Heun correction (sample.py:330-334)
        x_N, dxdt = out_and_grad_fn(net, x_next.to(F32), torch.full((B,), s_next, device=dev, dtype=F32), labels)
        x_N, dxdt = x_N.to(F64), dxdt.to(F64)
        d_prime = (x_next - x_N) / s_next
        x_next = x_cur + (s_next - s_cur) * (0.5 * d_cur + 0.5 * d_prime)
    loss_a, loss_u = obs_losses(x_N, obs_a, obs_u, mask_a, mask_u, ch_a)
    loss_pde = loss_fn(x_N[:, ch_a:], dxdt[:, ch_a:], labels, **loss_kwargs)
    w_a, w_u, w_pde = guidance_weights(i, num_steps, zeta_a, zeta_u, zeta_pde)
    loss_comb = w_a * loss_a + w_u * loss_u + w_pde * loss_pde
    internals = None
    if return_internals:                                         # seed gradient d loss / d x0-hat
        seeds = torch.autograd.grad(loss_comb, [x_N] + ([dxdt] if dxdt.requires_grad else []),
                                    retain_graph=True, allow_unused=True)
        internals = {"x_N": x_N.detach(), "dxdt": dxdt.detach(), "seed_x": seeds[0].detach(),
                     "seed_dxdt": seeds[1].detach() if len(seeds) > 1 and seeds[1] is not None else None}
    grad_x = torch.autograd.grad(loss_comb, x_cur)[0]
    x_next = (x_next - grad_x).detach()
    row = [float(loss_a.item()), float(loss_u.item()), float(loss_pde.item()), float(loss_comb.item())]
    if return_internals:
        internals["grad_x"] = grad_x.detach()
        return x_next, row, internals
    return x_next, row


def joint_sample(net, device, sample_shape, num_channels, ch_a, loss_fn, loss_kwargs,
                 labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde,
                 num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7.0,
                 out_and_grad_fn=X_and_dXdt_fd, num_samples=None, latents=None, return_trajectory=False):
    """Restatement of JointSampler.sample (sample.py:278-363).  ``latents`` (B,C,H,W) fp64 replaces
    the reference's ``torch.randn`` draw (sample.py:314) so CPU and CUDA runs share a start."""
    obs_u, mask_u = obs_u.to(device=device, dtype=F64), mask_u.to(device=device, dtype=F64)
    obs_a, mask_a = obs_a.to(device=device, dtype=F64), mask_a.to(device=device, dtype=F64)
    sigmas = karras_sigmas(num_steps, sigma_min, sigma_max, rho, device, net)
    B = labels.shape[0] if labels is not None else num_samples
    if labels is not None:
        labels = labels.to(device=device, dtype=F32)
    if latents is None:
        latents = torch.randn((B, num_channels, *sample_shape), device=device, dtype=F64)
    x_next = latents.to(device=device, dtype=F64) * sigmas[0]
    losses = torch.zeros((num_steps, 4))
    traj = []
    for i in range(num_steps):
        x_next, row = guided_step(net, x_next, i, sigmas, labels, obs_a, obs_u, mask_a, mask_u, ch_a,
                                  loss_fn, loss_kwargs, zeta_a, zeta_u, zeta_pde, num_steps, out_and_grad_fn)
        losses[i] = torch.tensor(row)
        if return_trajectory:
            traj.append(x_next.to(F32).cpu())
    x = x_next.to(F32).detach().cpu()
    if return_trajectory:
        return x, losses.numpy(), traj
    return x, losses.numpy()


def unconditional_sample(net, device, sample_shape, num_channels, labels=None, net_obs=None, num_steps=18, sigma_min=0.002,
                         sigma_max=80.0, rho=7.0, num_samples=None, latents=None):
    """Restatement of UnconditionalSampler.sample (sample.py:191-239): plain EDM Heun sampling, fp64 state and
    schedule, fp32 denoiser.  ``latents`` replaces the reference's ``torch.randn`` draw (sample.py:222)."""
    sigmas = karras_sigmas(num_steps, sigma_min, sigma_max, rho, device, net)
    B = labels.shape[0] if labels is not None else num_samples
    if labels is not None:
        labels = labels.to(device=device, dtype=F32)
    args = (labels,) if net_obs is None else (labels, net_obs.to(device=device, dtype=F32))
    if latents is None:
        latents = torch.randn((B, num_channels, *sample_shape), device=device, dtype=F64)
    x_next = latents.to(device=device, dtype=F64) * sigmas[0]
    with torch.no_grad():
        for i in range(num_steps):
            s_cur, s_next = sigmas[i], sigmas[i + 1]
            x_cur = x_next
            x_N = net(x_cur.to(F32), torch.full((B,), s_cur, device=device, dtype=F32), *args).to(F64)
            d_cur = (x_cur - x_N) / s_cur
            x_next = x_cur + (s_next - s_cur) * d_cur
            if i < num_steps - 1:
                x_N = net(x_next.to(F32), torch.full((B,), s_next, device=device, dtype=F32), *args).to(F64)
                d_prime = (x_next - x_N) / s_next
                x_next = x_cur + (s_next - s_cur) * (0.5 * d_cur + 0.5 * d_prime)
    return x_next.to(F32).detach().cpu()


def edm_heat_loss(net, x, labels, dx, noise, pde_loss_coeff=1.0, method="joint", residual_estimation="ME", P_mean=-1.2,
                  P_std=1.2, sigma_data=0.5, reduce_method="mean", sigma_min=0.01, rho=7.0, steps=2):
    """Restatement of EDMHeatLoss.__call__ (models/loss.py:127-171) with the two torch.randn draws (loss.py:129,132)
    passed in as ``noise = (rnd_normal, eps)``.  Same dtype as the reference: the Laplacian is an fp32 conv2d."""
    ch_a = 1 if method == "joint" else 0
    rnd_normal, eps = noise
    sigma = (rnd_normal * P_std + P_mean).exp()
    weight = (sigma ** 2 + sigma_data ** 2) / (sigma * sigma_data) ** 2
    n = eps * sigma
    D_yn, dxdt = X_and_dXdt_fd(net, x + n, sigma.flatten(), labels, no_grad=False)
    dxdt = dxdt.detach()[:, ch_a:, ...]
    edm_loss = weight * ((D_yn - x) ** 2)
    if residual_estimation == "ME":
        x_0star = D_yn
    else:   # two_step_sample, loss.py:78-124
        B = x.shape[0]
        s_max = sigma.view(B)
        s_min = torch.tensor(float(sigma_min), device=x.device, dtype=torch.float32)
        idx = torch.arange(steps + 1, dtype=torch.float32, device=x.device)
        sig = torch.stack([(s_max[i] ** (1.0 / rho) + idx / steps * (s_min ** (1.0 / rho) - s_max[i] ** (1.0 / rho))) ** rho
                           for i in range(B)], dim=0)
        x_0star = D_yn
        for s_cur, s_next in zip(sig.T[:-1], sig.T[1:]):
            x_N = net(x_0star, s_cur.flatten(), labels)
            x_0star = x_0star + (s_next.view(B, 1, 1, 1) - s_cur.view(B, 1, 1, 1)) * ((x_0star - x_N) / s_cur.view(B, 1, 1, 1))
    pde_loss = (dxdt - labels[:, 1].view(-1, 1, 1, 1) * laplacian(x_0star[:, ch_a:, ...], dx)) ** 2 / (x.shape[-2] * x.shape[-1])
    if reduce_method == "mean":
        edm_loss = edm_loss.mean(dim=(1, 2, 3))
        pde_loss = pde_loss.mean(dim=(1, 2, 3)) * pde_loss_coeff / (sigma ** 2)
    else:
        edm_loss = edm_loss.sum(dim=(1, 2, 3))
        pde_loss = pde_loss.sum(dim=(1, 2, 3)) * pde_loss_coeff / (sigma ** 2)
    return edm_loss + pde_loss


# ----------------------------------------------------------------------------------------
# closed-form seed gradients (numpy, fp64) -- what the CUDA VJP kernels implement
# ----------------------------------------------------------------------------------------
def heat_guidance_numpy(x_N, dxdt, alpha, dx, obs_a, obs_u, mask_a, mask_u, ch_a, w_a, w_u, w_pde):
    """Losses and analytic d loss_comb / d (x_N, dxdt_u) for the heat path; all arrays fp64.

    d loss_pde / d u = -(alpha/dx^2) K^T r / (H W loss_pde),  d / d dudt = r / (H W loss_pde),
    d loss_obs / d x = mask^2 (x - obs) / loss_obs  (SURVEY.md section 8 rows a-3, a-6).
    """
    x_N = np.asarray(x_N, np.float64)
    B, C, H, W = x_N.shape
    a, u = x_N[:, :ch_a], x_N[:, ch_a:]
    al = np.asarray(alpha, np.float64).reshape(B, 1, 1, 1)
    r = np.asarray(dxdt, np.float64)[:, ch_a:] - al * laplacian_numpy(u, dx)
    loss_pde = math.sqrt(float((r ** 2).sum()) / (H * W))
    g = np.zeros_like(x_N)
    g_dudt = w_pde * r / (H * W * loss_pde)
    g[:, ch_a:] += -w_pde * al * laplacian_adjoint_numpy(r, dx) / (H * W * loss_pde)
    loss_a = loss_u = 0.0
    ma = np.broadcast_to(np.asarray(mask_a, np.float64), a.shape)
    mu = np.broadcast_to(np.asarray(mask_u, np.float64), u.shape)
    if np.asarray(mask_a).sum() > 0:
        da = ma * (a - obs_a)
        loss_a = math.sqrt(float((da ** 2).sum()))
        g[:, :ch_a] += w_a * ma * da / loss_a
    if np.asarray(mask_u).sum() > 0:
        du = mu * (u - obs_u)
        loss_u = math.sqrt(float((du ** 2).sum()))
        g[:, ch_a:] += w_u * mu * du / loss_u
    loss_comb = w_a * loss_a + w_u * loss_u + w_pde * loss_pde
    return (loss_a, loss_u, loss_pde, loss_comb), g, g_dudt


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                     a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def llg_norm_guidance_numpy(m, w_pde=1.0):
    """llg_loss2 and its closed-form gradient: -(1-n) m/n / (sqrt(sum (1-n)^2) H W)."""
    m = np.asarray(m, np.float64)
    H, W = m.shape[-2:]
    n = np.sqrt((m ** 2).sum(axis=1, keepdims=True))
    root = math.sqrt(float(((1 - n) ** 2).sum()))
    return root / (H * W), -w_pde * (1 - n) * (m / n) / (root * H * W)


def llg_residual_guidance_numpy(m, dmdt, field_mT, dx, consts: LLGConstants = LLGConstants(), w_pde=1.0):
    """Residual loss sqrt(sum r^2)/(H W) and closed-form gradients w.r.t. m and dmdt (fp64 numpy).

    With seed g = d loss / d r, q = g x m, a = m x H:
      G_m = -gamma (H x g) - alpha (a x g + H x q),   G_H = -gamma q - alpha (q x m),
      d loss / d m = -tau [ G_m + (c_ex/dx^2) K^T G_H + c_an (e . G_H) e ],   d loss / d dmdt = g
    (checked against torch autograd in tests/test_oracle_llg.py).
    """
    m = np.asarray(m, np.float64)
    dmdt = np.asarray(dmdt, np.float64)
    B, _, H, W = m.shape
    h_ext = np.asarray(field_mT, np.float64).reshape(B, 3, 1, 1) / (1000 * consts.mu0)
    Hf = h_ext + consts.c_ex * laplacian_numpy(m, dx)
    e = np.asarray(consts.easy_axis, np.float64).reshape(1, 3, 1, 1)
    if consts.K0 != 0.0:
        Hf = Hf + consts.c_an * (m * e).sum(axis=1, keepdims=True) * e
    a = _cross(m, Hf)
    rhs = -consts.gamma * a - consts.alpha * _cross(m, a)
    r = dmdt - rhs * consts.tau
    root = math.sqrt(float((r ** 2).sum()))
    loss = root / (H * W)
    g = w_pde * r / (root * H * W)
    q = _cross(g, m)
    G_m = -consts.gamma * _cross(Hf, g) - consts.alpha * (_cross(a, g) + _cross(Hf, q))
    G_H = -consts.gamma * q - consts.alpha * _cross(q, m)
    gm = G_m + consts.c_ex * laplacian_adjoint_numpy(G_H, dx)
    if consts.K0 != 0.0:
        gm = gm + consts.c_an * (G_H * e).sum(axis=1, keepdims=True) * e
    return loss, -consts.tau * gm, g


def X_and_dXdt(net, x, sigma, labels):
    """Forward-mode time derivative in labels[:, 0] (sample.py:69-103)."""
    t0 = labels[:, 0]

    def f(t):
        lbl = labels.clone()
        lbl[:, 0] = t
        return net(x, sigma, lbl)

    return torch.func.jvp(f, (t0,), (torch.ones_like(t0),))
