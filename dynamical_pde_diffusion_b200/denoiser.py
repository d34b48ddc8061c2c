"""Denoisers used on either side of the guidance kernels (plain PyTorch, by design).

north_star keeps the U-Net forward/backward in PyTorch; the sampler accepts any callable
``net(x_f32, sigma_(B,), labels) -> x0hat``.  This module provides

* :class:`EDMUNet` / :class:`EDMPrecond` -- an EDM-preconditioned residual U-Net whose module
  tree and parameter names follow the reference network (``src/diffusion_pde/models/nets.py``:
  ``EDMUNet`` :217-340, ``ResBlock`` :153-211, ``EDMWrapper`` :343-366) so reference checkpoints
  load with ``load_state_dict`` unchanged (checked by ``tests/test_denoiser.py`` against the live
  reference and the golden fixture);
* :class:`PointwiseDenoiser` -- an analytic stand-in for grids where a U-Net cannot run
  (config 5, 4096^2: GroupNorm is a global spatial statistic and does not slab-decompose).

Nothing here launches our CUDA kernels; it exists so benchmarks and parity tests have a
denoiser on the GPU box, where the reference source is absent.
"""
from __future__ import annotations

import math

import torch
from torch import nn


def _kaiming_linear_(module: nn.Module) -> None:
    """fan-in Kaiming normal with gain 1 and zero bias (nets.py:6-13 with its defaults)."""
    nn.init.kaiming_normal_(module.weight, a=0, mode="fan_in", nonlinearity="linear")
    if module.bias is not None:
        nn.init.zeros_(module.bias)


def _zeros_(module: nn.Module) -> None:
    nn.init.zeros_(module.weight)
    if module.bias is not None:
        nn.init.zeros_(module.bias)


def _conv(cin: int, cout: int, k: int, *, up: bool = False, down: bool = False, zero: bool = False) -> nn.Module:
    """3x3 / 1x1 convolution with reflect padding; stride-2 conv down, transposed conv up (nets.py:133-150)."""
    pad = max(0, (k - 1) // 2)
    if up:
        layer = nn.ConvTranspose2d(cin, cout, k, stride=2, padding=pad, output_padding=1)
    else:
        layer = nn.Conv2d(cin, cout, k, stride=2 if down else 1, padding=pad, padding_mode="reflect")
    (_zeros_ if zero else _kaiming_linear_)(layer)
    return layer


def _groups(ch: int) -> int:
    return 32 if ch >= 32 and ch % 32 == 0 else ch


class PositionalEmbedding(nn.Module):
    """cos/sin features of the noise level at geometrically spaced frequencies (nets.py:29-42)."""

    def __init__(self, num_channels: int, max_positions: int = 10000, endpoint: bool = False):
        super().__init__()
        self.num_channels, self.max_positions, self.endpoint = num_channels, max_positions, endpoint

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        f = torch.arange(half, dtype=torch.float32, device=x.device) / (half - (1 if self.endpoint else 0))
        f = (1.0 / self.max_positions) ** f
        ang = torch.outer(x, f.to(x.dtype))
        return torch.cat([ang.cos(), ang.sin()], dim=1)


class ResBlock(nn.Module):
    """GroupNorm-SiLU-conv, + embedding, GroupNorm-SiLU-(zero-init conv), + skip, * 2^-1/2 (nets.py:153-211)."""

    def __init__(self, in_ch: int, out_ch: int, emb_ch: int, up: bool = False, down: bool = False,
                 dropout: float = 0.0, skip_scale: float = 2 ** -0.5):
        super().__init__()
        self.in_channels, self.out_channels = in_ch, out_ch
        self.skip_scale = skip_scale
        self.norm1 = nn.GroupNorm(_groups(in_ch), in_ch)
        self.norm2 = nn.GroupNorm(_groups(out_ch), out_ch)
        self.act = nn.SiLU()
        self.conv1 = _conv(in_ch, out_ch, 3, up=up, down=down)
        self.conv2 = _conv(out_ch, out_ch, 3, zero=True)
        self.emb_layer = nn.Linear(emb_ch, out_ch)
        _kaiming_linear_(self.emb_layer)
        self.skip = _conv(in_ch, out_ch, 1, up=up, down=down) if (in_ch != out_ch or up or down) else None
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
        h = self.conv1(self.act(self.norm1(x)))
        h = h + self.emb_layer(emb)[:, :, None, None]
        h = self.conv2(self.dropout(self.act(self.norm2(h))))
        h = h + (x if self.skip is None else self.skip(x))
        return h * self.skip_scale


class EDMUNet(nn.Module):
    """Encoder/decoder of ResBlocks with concatenated skips; sigma (+ label) embedding into every block."""

    def __init__(self, img_channels: int, obs_channels: int = 0, label_dim: int = 0, base_channels: int = 64,
                 channel_mults=(1, 2, 2), num_res_blocks: int = 2, dropout: float = 0.0,
                 sigma_emb_dim: int = 64, emb_dim: int = 256):
        super().__init__()
        self.img_channels, self.obs_channels = img_channels, obs_channels
        self.sigma_embed = PositionalEmbedding(sigma_emb_dim)
        self.time_mlp = nn.Sequential(nn.Linear(sigma_emb_dim, emb_dim), nn.SiLU(), nn.Linear(emb_dim, emb_dim))
        _kaiming_linear_(self.time_mlp[0])
        _kaiming_linear_(self.time_mlp[2])
        self.label_embed = nn.Linear(label_dim, emb_dim) if label_dim > 0 else None
        if self.label_embed is not None:
            _kaiming_linear_(self.label_embed)

        widths = [base_channels * m for m in channel_mults]
        self.enc = nn.ModuleList()
        skip_ch = []
        ch = widths[0]
        for lvl, w in enumerate(widths):
            self.enc.append(_conv(img_channels + obs_channels, w, 3) if lvl == 0
                            else ResBlock(ch, w, emb_dim, down=True, dropout=dropout))
            skip_ch.append(w)
            for _ in range(num_res_blocks):
                self.enc.append(ResBlock(w, w, emb_dim, dropout=dropout))
                skip_ch.append(w)
            ch = w

        self.dec = nn.ModuleList()
        for lvl in range(len(widths) - 1, -1, -1):
            if lvl == len(widths) - 1:
                self.dec += [ResBlock(ch, ch, emb_dim, dropout=dropout), ResBlock(ch, ch, emb_dim, dropout=dropout)]
            else:
                self.dec.append(ResBlock(ch, ch, emb_dim, up=True, dropout=dropout))
            for _ in range(num_res_blocks + 1):
                self.dec.append(ResBlock(ch + skip_ch.pop(), widths[lvl], emb_dim, dropout=dropout))
                ch = widths[lvl]
        self.final_block = nn.Sequential(nn.GroupNorm(32 if ch % 32 == 0 else ch, ch),
                                         _conv(ch, img_channels, 3, zero=True))

    def forward(self, x, sigma, labels=None, obs=None):
        if obs is not None and self.obs_channels > 0:
            x = torch.cat([x, obs], dim=1)
        emb = self.time_mlp(self.sigma_embed(sigma))
        if self.label_embed is not None and labels is not None:
            emb = emb + self.label_embed(labels)
        skips = []
        for blk in self.enc:
            x = blk(x, emb) if isinstance(blk, ResBlock) else blk(x)
            skips.append(x)
        for blk in self.dec:
            if x.shape[1] != blk.in_channels:
                x = torch.cat([x, skips.pop()], dim=1)
            x = blk(x, emb)
        return self.final_block(x)


class EDMPrecond(nn.Module):
    """D(x; sigma) = c_skip x + c_out F(c_in x; ln(sigma)/4) (nets.py:343-366).  The attribute is named
    ``unet`` so ``EDMWrapper`` checkpoints load directly."""

    def __init__(self, unet: nn.Module, sigma_data: float = 0.5):
        super().__init__()
        self.unet, self.sigma_data = unet, sigma_data

    def forward(self, x, sigma, *args, **kwargs):
        s = torch.reshape(sigma, (-1, 1, 1, 1))
        sd = self.sigma_data
        c_skip = sd ** 2 / (s ** 2 + sd ** 2)
        c_out = s * sd / torch.sqrt(s ** 2 + sd ** 2)
        c_in = 1 / torch.sqrt(s ** 2 + sd ** 2)
        c_noise = torch.flatten(torch.log(s) / 4).to(torch.float32)
        return c_skip * x + c_out * self.unet(c_in * x, c_noise, *args, **kwargs)


EDMWrapper = EDMPrecond  # reference spelling


def build_unet_v2(img_channels: int, label_dim: int, **overrides) -> EDMPrecond:
    """The ``conf/model/unetv2.yaml`` network: base 64, mults (1,2,2), 2 res blocks, emb 256, sigma_data 0.5."""
    cfg = dict(base_channels=64, channel_mults=(1, 2, 2), num_res_blocks=2, dropout=0.0, sigma_emb_dim=64, emb_dim=256)
    cfg.update(overrides)
    sigma_data = cfg.pop("sigma_data", 0.5)
    return EDMPrecond(EDMUNet(img_channels=img_channels, label_dim=label_dim, **cfg), sigma_data=sigma_data)


@torch.no_grad()
def randomize_zero_init(net: nn.Module, seed: int = 0, scale: float = 1.0) -> nn.Module:
    """Give the zero-initialised convolutions seeded Kaiming weights.

    An untrained reference network has ``conv2`` and the output conv at zero (nets.py:181,300), so
    D(x) = c_skip x and its Jacobian is trivial; synthetic benchmarks and parity tests need a denoiser
    with a non-trivial Jacobian.
    """
    gen = torch.Generator().manual_seed(seed)
    for mod in net.modules():
        if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)) and float(mod.weight.abs().sum()) == 0.0:
            fan_in = mod.weight[0].numel() if isinstance(mod, nn.Conv2d) else mod.weight[:, 0].numel()
            w = torch.randn(mod.weight.shape, generator=gen) * (scale / math.sqrt(fan_in))
            mod.weight.copy_(w.to(mod.weight))
    return net


class PointwiseDenoiser(nn.Module):
    """``D(x; sigma) = c_skip(sigma) x + c_out(sigma) tanh(c_in(sigma) x + t)`` with EDM preconditioning coefficients
    (``models/nets.py:352-366``) and the label's time entry as a bias: local, differentiable, label dependent (so the
    finite-difference time derivative is not identically zero).  A stand-in for throughput and parity runs on grids
    where the reference U-Net cannot run (config 5, 4096^2: activations exceed HBM and GroupNorm is a global
    statistic, ``models/nets.py:172-175``); not a trained model.  No spatial coupling, no parameters."""

    def __init__(self, sigma_data: float = 0.5):
        super().__init__()
        self.sigma_data = sigma_data

    def forward(self, x, sigma, labels=None, **kw):
        s = sigma.to(x.dtype).reshape(-1, 1, 1, 1)
        sd = self.sigma_data
        c_skip = sd ** 2 / (s ** 2 + sd ** 2)
        c_out = s * sd / (s ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + s ** 2).sqrt()
        t = labels[:, 0].to(x.dtype).reshape(-1, 1, 1, 1) if labels is not None else 0.0
        return c_skip * x + c_out * torch.tanh(c_in * x + t)

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)
