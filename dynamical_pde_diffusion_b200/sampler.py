"""Drop-in ``JointSampler`` (Level 2 of the boundary).

Same constructor and ``sample()`` signature and return values as the reference class
(``src/diffusion_pde/sampling/sample.py:243-363``); the attributes ``sampling_context`` and ``test_loop`` read
(``net``, ``device``, ``num_channels``, ``sample_shape``, ``num_samples``; ``sample.py:626-636``,
``model_testing.py:175-196``) are kept.  Per guided step the reference's ~40 ATen ops + autograd mirror + 6 host
syncs become:

    denoiser (PyTorch)  ->  dpde_euler_predict  ->  denoiser (PyTorch)
    ->  dpde_guidance_reduce + dpde_guidance_vjp   (losses, analytic seed gradient d loss_comb / d x0-hat)
    ->  torch.autograd.grad through the denoiser(s) only (dpde_euler_predict_bwd links the two evaluations)
    ->  dpde_heun_guided_update                    (Heun average + guidance update, fp64 state + fp32 copy)

with no host synchronisation inside the loop (the loss trace stays on the device until ``sample()`` returns).
Additive keyword arguments only: ``latents=`` / ``generator=`` (reproducible starts), ``group=`` / ``coupled=``
(multi-GPU, see ``distributed.py``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _ffi
from .ops import GuidanceEngine, LLGConstants, _stream
from ._ffi import PDE_HEAT, PDE_LLG_NORM, PDE_LLG_RESIDUAL, PDE_NONE

F32, F64 = torch.float32, torch.float64


# ---------------------------------------------------------------------------------------------------------
# denoiser + time-derivative providers (sample.py:15-103): they only call the net, so they stay PyTorch
# ---------------------------------------------------------------------------------------------------------
def X_and_dXdt_dummy(net, x, sigma, labels, **kwargs):
    """Denoised estimate with a zero time derivative (``sample.py:15-18``)."""
    out = net(x, sigma, labels, **kwargs)
    return out, torch.zeros_like(out)


def X_and_dXdt_fd(net, x, sigma, labels, eps=1e-5, no_grad=True, **kwargs):
    """Central finite difference in ``labels[:, 0]`` (time); the two offset evaluations carry no graph unless
    ``no_grad=False`` (``sample.py:21-66``)."""
    if labels is None:
        return X_and_dXdt_dummy(net, x, sigma, labels, **kwargs)
    plus, minus = labels.detach().clone(), labels.detach().clone()
    plus[:, 0] += eps
    minus[:, 0] -= eps
    with torch.no_grad() if no_grad else torch.enable_grad():
        up = net(x, sigma, plus, **kwargs)
        um = net(x, sigma, minus, **kwargs)
    dudt = (up - um) / (2 * eps)
    return net(x, sigma, labels, **kwargs), dudt


def X_and_dXdt_fd_batched(net, x, sigma, labels, eps=1e-5, **kwargs):
    """``X_and_dXdt_fd`` with the two offset evaluations batched into ONE denoiser call of batch 2B: two
    thirds of the finite-difference launches disappear.  Per-sample arithmetic is unchanged, but a batch of 2B may make
    cuDNN pick other kernels, and the difference quotient amplifies their 1e-7 differences by 1/(2 eps) -- so this is an
    opt-in provider, not the parity default."""
    if labels is None:
        return X_and_dXdt_dummy(net, x, sigma, labels, **kwargs)
    with torch.no_grad():
        lbl = labels.detach().repeat(2, 1)
        lbl[: labels.shape[0], 0] += eps
        lbl[labels.shape[0]:, 0] -= eps
        xd = x.detach()
        out = net(torch.cat([xd, xd]), torch.cat([sigma, sigma]), lbl, **kwargs)
        up, um = out[: labels.shape[0]], out[labels.shape[0]:]
        dudt = torch.sub(up, um).div_(2 * eps)
    return net(x, sigma, labels, **kwargs), dudt


def X_and_dXdt(net, x, sigma, labels):
    """Exact time derivative by forward-mode AD in ``labels[:, 0]`` (``sample.py:69-103``)."""
    t0 = labels[:, 0]

    def f(t):
        lbl = labels.clone()
        lbl[:, 0] = t
        return net(x, sigma, lbl)

    return torch.func.jvp(f, (t0,), (torch.ones_like(t0),))


class sampling_context:
    """``sampling_context`` of the reference (``sample.py:622-637``): cuDNN convolutions in TF32, the denoiser in
    ``eval()`` mode and on the sampler's device while the block runs.  On exit the precision setting is restored and --
    as the reference does -- the net goes back to the CPU (``keep_on_device=True`` skips that move and the cache flush,
    for callers that sample repeatedly).  ``tf32=False`` leaves the precision alone (parity runs use IEEE fp32)."""

    def __init__(self, sampler, tf32: bool = True, keep_on_device: bool = False):
        self.sampler, self.tf32, self.keep = sampler, tf32, keep_on_device

    def __enter__(self):
        self.prev = torch.backends.cudnn.conv.fp32_precision
        if self.tf32:
            torch.backends.cudnn.conv.fp32_precision = "tf32"
        self.was_training = self.sampler.net.training
        self.sampler.net.eval()
        self.sampler.net.to(self.sampler.device)
        return self.sampler

    def __exit__(self, exc_type, exc_value, traceback):
        torch.backends.cudnn.conv.fp32_precision = self.prev
        if not self.keep:
            self.sampler.net.to(torch.device("cpu"))
            torch.cuda.empty_cache()
        torch.cuda.synchronize()


class _EulerPredict(torch.autograd.Function):
    """x_eu32 = fp32(x_cur + h (x_cur - x0)/s_cur) as a graph node between the two denoiser evaluations.

    Only ``x0`` (the first denoiser output) is a differentiable input; the state's own contribution to the
    gradient is added analytically by ``dpde_heun_guided_update``, which reads the incoming gradient this node
    stashes -- exactly the three terms autograd sums in the reference (``sample.py:327-328,354``).
    """

    @staticmethod
    def forward(ctx, x0, x_cur64, s_cur, s_next, stash):
        out = torch.empty_like(x0)
        _ffi.call("dpde_euler_predict", x_cur64.data_ptr(), x0.data_ptr(), s_cur, s_next, out.data_ptr(), x0.numel(), _stream())
        ctx.s_cur, ctx.s_next, ctx.stash = s_cur, s_next, stash
        return out

    @staticmethod
    def backward(ctx, g_eu):
        g_eu = g_eu.contiguous()
        ctx.stash["g_eu"] = g_eu
        seed = torch.empty_like(g_eu)
        _ffi.call("dpde_euler_predict_bwd", g_eu.data_ptr(), ctx.s_cur, ctx.s_next, seed.data_ptr(), g_eu.numel(), _stream())
        return seed, None, None, None, None


def _pde_kind_of(loss_fn):
    kind = getattr(loss_fn, "_dpde_kind", None)
    if kind is not None:
        return kind
    # the reference's own functions, when a caller passes them unchanged
    name, mod = getattr(loss_fn, "__name__", ""), getattr(loss_fn, "__module__", "") or ""
    if mod.endswith("pde_losses"):
        return {"heat_loss2": PDE_HEAT, "llg_loss2": PDE_LLG_NORM}.get(name)
    return None


def _f32c(t):
    t = t if t.dtype == F32 else t.to(F32)
    return t if t.is_contiguous() else t.contiguous()


class Sampler:
    """Base class, as in the reference (``sample.py:137-143``)."""

    # ---- schedule (sample.py:206-209, 305-308) -----------------------------------------------------------
    def _sigmas(self, num_steps, sigma_min, sigma_max, rho):
        idx = torch.arange(num_steps, dtype=F64, device=self.device)
        s = (sigma_max ** (1.0 / rho) + idx / (num_steps - 1) * (sigma_min ** (1.0 / rho) - sigma_max ** (1.0 / rho))) ** rho
        s = getattr(self.net, "round_sigma", lambda v: v)(s)
        return torch.cat([s, torch.zeros_like(s[:1])]).tolist()      # one D2H per sample() call


class UnconditionalSampler(Sampler):
    """EDM Heun sampler without guidance, with the reference's API (``sample.py:145-239``), on the same kernels as
    the guided sampler: ``dpde_sampler_init`` (latents * sigma_0 + the fp32 copy the denoiser reads),
    ``dpde_euler_predict`` (the fp32 Euler state of the second evaluation) and ``dpde_heun_guided_update`` with no
    gradient operands.  The fp64 state never leaves the device and nothing synchronises the host inside the loop."""

    def __init__(self, net, device, sample_shape, num_channels, num_samples, num_steps=18, sigma_min=0.002, sigma_max=80.0,
                 rho=7.0):
        self.net = net
        self.device = device
        self.sample_shape = sample_shape
        self.num_channels = num_channels
        self.num_samples = num_samples
        self.num_steps = num_steps
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.rho = rho
        self.dtype_f = F32
        self.dtype_t = F64

    @torch.no_grad()
    def sample(self, labels=None, net_obs=None, num_steps=None, sigma_min=None, sigma_max=None, rho=None, *, latents=None,
               generator=None):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError(f"dpde_b200.UnconditionalSampler runs on CUDA devices only (got {dev}); there is no CPU path")
        num_steps = num_steps if num_steps is not None else self.num_steps
        sigma_min = sigma_min if sigma_min is not None else self.sigma_min
        sigma_max = sigma_max if sigma_max is not None else self.sigma_max
        rho = rho if rho is not None else self.rho
        sigmas = self._sigmas(num_steps, sigma_min, sigma_max, rho)
        B = labels.shape[0] if labels is not None else self.num_samples
        if labels is not None:
            labels = labels.to(device=dev, dtype=F32)
        args = (labels,)
        if net_obs is not None:
            args = (labels, net_obs.to(device=dev, dtype=F32))
        shape = (B, self.num_channels, *self.sample_shape)
        if latents is None:   # first RNG call, as sample.py:222
            latents = torch.randn(shape, device=dev, dtype=F64, generator=generator)
        else:
            latents = latents.to(device=dev, dtype=F64).contiguous()
        x64, x64n = torch.empty(shape, dtype=F64, device=dev), torch.empty(shape, dtype=F64, device=dev)
        x32, x_eu = torch.empty(shape, dtype=F32, device=dev), torch.empty(shape, dtype=F32, device=dev)
        n, st = x64.numel(), _stream()
        _ffi.call("dpde_sampler_init", latents.data_ptr(), sigmas[0], x64.data_ptr(), x32.data_ptr(), n, st)
        for i in range(num_steps):
            s_cur, s_next = sigmas[i], sigmas[i + 1]
            x0_1 = _f32c(self.net(x32, torch.full((B,), s_cur, device=dev, dtype=F32), *args))
            x0_2 = None
            if i < num_steps - 1:
                _ffi.call("dpde_euler_predict", x64.data_ptr(), x0_1.data_ptr(), s_cur, s_next, x_eu.data_ptr(), n, st)
                x0_2 = _f32c(self.net(x_eu, torch.full((B,), s_next, device=dev, dtype=F32), *args))
            x32n = torch.empty_like(x32)
            _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0_1.data_ptr(), x0_2.data_ptr() if x0_2 is not None else None,
                      None, None, s_cur, s_next, x64n.data_ptr(), x32n.data_ptr(), n, st)
            x64, x64n, x32 = x64n, x64, x32n
        return x32.detach().cpu()


class JointSampler(Sampler):
    """Physics-guided EDM Heun sampler with the reference's API (``sample.py:243-363``)."""

    def __init__(self, net, device, sample_shape, num_channels, num_samples, ch_a, loss_fn, loss_kwargs,
                 num_steps=18, sigma_min=0.002, sigma_max=80.0, rho=7.0, out_and_grad_fn=X_and_dXdt_fd,
                 *, group=None, coupled=False):
        self.net = net
        self.device = device
        self.sample_shape = sample_shape
        self.num_channels = num_channels
        self.num_samples = num_samples
        self.ch_a = ch_a
        self.loss_fn = loss_fn
        self.loss_kwargs = loss_kwargs
        self.num_steps = num_steps
        self.sigma_min = sigma_min
        self.sigma_max = sigma_max
        self.rho = rho
        self.out_and_grad_fun = out_and_grad_fn
        self.dtype_f = F32   # net runs in fp32
        self.dtype_t = F64   # state and schedule in fp64 (sample.py:275-276)
        self.group, self.coupled = group, coupled
        self._run = None

    def _allreduce(self):
        if not self.coupled:
            return None
        import torch.distributed as dist
        group = self.group
        return lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    # ---- one sample() call is split into begin / step / finish so benchmarks can time single steps ------
    def begin(self, labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, num_steps=None, sigma_min=None,
              sigma_max=None, rho=None, latents=None, generator=None):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError(f"dpde_b200.JointSampler runs on CUDA devices only (got {dev}); there is no CPU path")
        num_steps = num_steps if num_steps is not None else self.num_steps
        sigma_min = sigma_min if sigma_min is not None else self.sigma_min
        sigma_max = sigma_max if sigma_max is not None else self.sigma_max
        rho = rho if rho is not None else self.rho
        H, W = self.sample_shape
        C_, ch_a = self.num_channels, self.ch_a
        obs_u, mask_u = obs_u.to(dev, non_blocking=True), mask_u.to(dev, non_blocking=True)
        obs_a, mask_a = obs_a.to(dev, non_blocking=True), mask_a.to(dev, non_blocking=True)
        sigmas = self._sigmas(num_steps, sigma_min, sigma_max, rho)
        B = labels.shape[0] if labels is not None else self.num_samples
        if labels is not None:
            labels = labels.to(device=dev, dtype=F32)
        if latents is None:   # first RNG call, as sample.py:314
            latents = torch.randn((B, C_, H, W), device=dev, dtype=F64, generator=generator)
        else:
            latents = latents.to(device=dev, dtype=F64).contiguous()

        kind = _pde_kind_of(self.loss_fn)
        coef, dx, llg = None, 0.0, None
        if kind == PDE_HEAT:
            coef, dx = labels[:, -1].to(F64), float(self.loss_kwargs["dx"])
        elif kind == PDE_LLG_RESIDUAL:
            llg = self.loss_kwargs.get("consts", LLGConstants())
            coef, dx = labels[:, -3:].to(F64) / (1000 * llg.mu0), float(self.loss_kwargs["dx"])
        has_flags = None
        if self.coupled:
            # one batch-B run over all ranks: the empty-mask branches (sample.py:339,341) are decided on the GLOBAL batch
            # (a rank whose per-sample mask shard happens to be empty must still take part in the norm), and an arbitrary
            # loss_fn would be evaluated per shard -- its term cannot be coupled, so it is refused
            if kind is None:
                raise RuntimeError("JointSampler(coupled=True) needs one of the fused residuals (heat_loss2, llg_loss2, "
                                   "llg_residual_loss): an arbitrary loss_fn is evaluated per shard and cannot be coupled")
            import torch.distributed as dist
            flags = torch.tensor([float(mask_a.sum() > 0), float(mask_u.sum() > 0)], device=dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
            has_flags = (bool(flags[0] > 0), bool(flags[1] > 0))
        engine = GuidanceEngine(B, C_, ch_a, H, W, kind if kind is not None else PDE_NONE, dev, obs_a=obs_a, mask_a=mask_a,
                                obs_u=obs_u, mask_u=mask_u, sample_coef=coef, dx=dx, llg=llg, has_flags=has_flags)
        x64 = torch.empty((B, C_, H, W), dtype=F64, device=dev)
        x32 = torch.empty((B, C_, H, W), dtype=F32, device=dev)
        _ffi.call("dpde_sampler_init", latents.data_ptr(), sigmas[0], x64.data_ptr(), x32.data_ptr(), x64.numel(), _stream())
        self._run = dict(engine=engine, fused=kind is not None, sigmas=sigmas, N=num_steps, B=B, labels=labels, x64=x64,
                         x32=x32, x64_alt=torch.empty_like(x64), zetas=(zeta_a, zeta_u, zeta_pde),
                         trace=torch.zeros((num_steps, 4), dtype=F32, device=dev), i=0, allreduce=self._allreduce())
        return self._run

    def step(self):
        """One guided Heun step (``sample.py:320-357``); advances the internal state, no host sync."""
        ctx = self._step_front()
        self._step_back(ctx)

    def _step_front(self):
        """Denoiser evaluation(s), Euler predictor and pass 1 (the three sums).  With a cross-rank all-reduce
        (``coupled`` batch shards, row slabs) the sums are left un-finalised for the caller to reduce."""
        r = self._run
        i, N, B = r["i"], r["N"], r["B"]
        s_cur, s_next = r["sigmas"][i], r["sigmas"][i + 1]
        dev, labels, engine = r["x64"].device, r["labels"], r["engine"]
        last = not (i < N - 1)
        x32 = r["x32"].requires_grad_(True)
        stash = {}
        with torch.enable_grad():
            x0_1, dxdt_1 = self.out_and_grad_fun(self.net, x32, torch.full((B,), s_cur, device=dev, dtype=F32), labels)
            x0_1c = _f32c(x0_1)
            if not last:
                x_eu = _EulerPredict.apply(x0_1c, r["x64"], s_cur, s_next, stash)
                x0_2, dxdt_2 = self.out_and_grad_fun(self.net, x_eu, torch.full((B,), s_next, device=dev, dtype=F32), labels)
                xN, dxdt = _f32c(x0_2), dxdt_2
            else:
                xN, dxdt = x0_1c, dxdt_1
        # guidance weights (sample.py:348-351), Python floats
        za, zu, zp = r["zetas"]
        w = (za, zu, zp) if i <= 0.8 * N else (0.1 * za, 0.1 * zu, zp)
        dx_c = _f32c(dxdt.detach()) if dxdt is not None else None
        want_d = dxdt is not None and dxdt.requires_grad
        ctx = dict(x32=x32, stash=stash, x0_1c=x0_1c, xN=xN, dxdt=dxdt, dx_c=dx_c, want_d=want_d, w=w, last=last,
                   s_cur=s_cur, s_next=s_next)
        if r["fused"]:
            self._pass1(r, engine, xN.detach(), dx_c, w, i)
        return ctx

    # ---- the two halves of the sum reduction; multi-GPU subclasses (row slabs) replace them ----------------------
    def _pass1(self, r, engine, xN, dx_c, w, i):
        engine.reduce(xN, dx_c, w, trace_row=r["trace"][i] if r["allreduce"] is None else None, finalize=r["allreduce"] is None)

    def _combine(self, r, engine, i):
        if r["allreduce"] is not None:
            r["allreduce"](engine.sums)
            engine.finalize(r["trace"][i])

    def _step_back(self, ctx):
        """(all-reduce of the sums,) seed gradient, backward through the denoiser(s), fused Heun + guidance update."""
        r = self._run
        i, engine = r["i"], r["engine"]
        xN, dxdt, dx_c, want_d, w, last = ctx["xN"], ctx["dxdt"], ctx["dx_c"], ctx["want_d"], ctx["w"], ctx["last"]
        if r["fused"]:
            self._combine(r, engine, i)
            g, gd = engine.vjp(xN.detach(), dx_c, w, want_d)
        else:
            g, gd = self._seed_generic(engine, xN.detach(), dx_c, w, r["trace"][i], want_d, r["allreduce"])
        outs, seeds = [xN], [g]
        if want_d and gd is not None:
            outs.append(dxdt)
            seeds.append(gd.to(dxdt.dtype))
        (g_cur,) = torch.autograd.grad(outs, [ctx["x32"]], grad_outputs=seeds, allow_unused=True)
        g_eu = ctx["stash"].get("g_eu")
        x64n = r["x64_alt"]
        x32n = r["x32_alt"] if r.get("x32_alt") is not None else torch.empty_like(r["x32"])
        self._launch_update(r["x64"], ctx["x0_1c"], None if last else xN, g_eu, _f32c(g_cur) if g_cur is not None else None,
                            ctx["s_cur"], ctx["s_next"], x64n, x32n)
        x32_old = r["x32"].detach().requires_grad_(False) if r.get("x32_alt") is not None else None
        r["x64"], r["x64_alt"], r["x32"], r["i"] = x64n, r["x64"], x32n, i + 1
        if x32_old is not None:
            r["x32_alt"] = x32_old

    def _launch_update(self, x64, x0_1c, x0_2, g_eu, g_cur, s_cur, s_next, x64n, x32n):
        _ffi.call("dpde_heun_guided_update", x64.data_ptr(), x0_1c.data_ptr(), x0_2.data_ptr() if x0_2 is not None else None,
                  g_eu.data_ptr() if g_eu is not None else None, g_cur.data_ptr() if g_cur is not None else None,
                  s_cur, s_next, x64n.data_ptr(), x32n.data_ptr(), x64n.numel(), _stream())

    def _seed_generic(self, engine, xN, dxdt, w, trace_row, want_d, allreduce):
        """Arbitrary ``loss_fn`` plug-in: observation terms from the kernels, the PDE term through the callable
        (fp64, ``sample.py:345-347``) and torch autograd.  Everything is summed in fp64 and rounded to the
        denoiser's fp32 once, like the fused path, so both routes give the same seed."""
        ch_a = self.ch_a
        x64 = xN.to(F64)
        d64 = dxdt.to(F64) if dxdt is not None else None
        with torch.enable_grad():
            xv = x64.detach().requires_grad_(True)
            dv = d64.detach().requires_grad_(want_d) if d64 is not None else None
            loss_pde = self.loss_fn(xv[:, ch_a:], dv[:, ch_a:] if dv is not None else None, self._run["labels"], **self.loss_kwargs)
            loss_pde = loss_pde.reshape(())
            grads = torch.autograd.grad(w[2] * loss_pde, [xv] + ([dv] if want_d else []), allow_unused=True)
        g, _ = engine.seed(x64, d64, (w[0], w[1], 0.0), trace_row=trace_row, want_dxdt_grad=False, allreduce=allreduce)
        if grads[0] is not None:
            g = g + grads[0]
        trace_row[2] = loss_pde.detach().to(F32)
        trace_row[3] = (engine.scalars[3] + w[2] * loss_pde.detach()).to(F32)
        gd = grads[1].to(F32) if want_d and len(grads) > 1 and grads[1] is not None else None
        return g.to(F32), gd

    def finish(self, return_losses=False, to_cpu=True):
        r = self._run
        x = r["x32"].detach()                                         # fp32 at sigma = 0 (sample.py:360)
        if to_cpu:                                                    # the reference always returns a CPU tensor
            x = x.cpu()
        losses = r["trace"].cpu().numpy() if return_losses else None  # (N,4) numpy (sample.py:362)
        self._run = None
        return x, losses

    def sample(self, labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, return_losses=False,
               num_steps=None, sigma_min=None, sigma_max=None, rho=None, *, latents=None, generator=None, to_cpu=True):
        """``sample.py:278-363``.  Additive keywords: ``latents`` / ``generator`` replace the ``torch.randn`` draw,
        ``to_cpu=False`` leaves the samples on the device (``evaluation.test_loop`` computes its metrics there)."""
        run = self.begin(labels, obs_a, obs_u, mask_a, mask_u, zeta_a, zeta_u, zeta_pde, num_steps, sigma_min, sigma_max,
                         rho, latents, generator)
        for _ in range(run["N"]):
            self.step()
        return self.finish(return_losses, to_cpu)
