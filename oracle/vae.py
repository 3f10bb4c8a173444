"""Oracle AutoencoderKL decoder (test infrastructure).

fp32 PyTorch restatement of `vae.decode(latents / scaling_factor)` as run by the
diffusers pipeline behind `backends/cuda_worker.py:222` (mirrored at
`backends/rknnlcm.py:614-618`).  Architecture per SURVEY.md Appendix A.3/A.4;
diffusers state-dict key names (`post_quant_conv.*`, `decoder.*`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import ResnetBlock2D, Upsample2D


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    sample_size: int = 512

    @staticmethod
    def tiny() -> "VAEConfig":
        return VAEConfig(block_out_channels=(64, 64, 128, 128), norm_num_groups=32, sample_size=128)


class VAEAttention(nn.Module):
    """diffusers `Attention(heads=1, bias=True, residual_connection=True, norm=GroupNorm)`."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1)
        o = self.to_out[0](att @ v)
        return x + o.transpose(1, 2).reshape(B, C, H, W)


class VAEMid(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, 1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([VAEAttention(c, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class VAEUpBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, up):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, 1e-6) for i in range(n)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = cfg.block_out_channels
        g = cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[-1], 3, padding=1)
        self.mid_block = VAEMid(ch[-1], g)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        prev = rev[0]
        for i, c in enumerate(rev):
            self.up_blocks.append(VAEUpBlock(prev, c, cfg.layers_per_block + 1, g,
                                             up=i != len(ch) - 1))
            prev = c
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class OracleVAEDecoder(nn.Module):
    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)

    def decode(self, z):
        """`AutoencoderKL.decode` (untiled): z is already divided by scaling_factor."""
        return self.decoder(self.post_quant_conv(z))

    # --- `vae.enable_tiling()` (`backends/cuda_worker.py:91`), SURVEY.md App. A.4 ---
    def tiled_decode(self, z):
        tile_latent = self.cfg.sample_size // 8          # tile_latent_min_size
        overlap = int(tile_latent * (1 - 0.25))
        blend = int(self.cfg.sample_size * 0.25)
        limit = self.cfg.sample_size - blend
        rows = []
        for i in range(0, z.shape[2], overlap):
            row = []
            for j in range(0, z.shape[3], overlap):
                row.append(self.decode(z[:, :, i:i + tile_latent, j:j + tile_latent]))
            rows.append(row)
        out_rows = []
        for i, row in enumerate(rows):
            out_row = []
            for j, tile in enumerate(row):
                if i > 0:
                    tile = _blend_v(rows[i - 1][j], tile, blend)
                if j > 0:
                    tile = _blend_h(row[j - 1], tile, blend)
                out_row.append(tile[:, :, :limit, :limit])
            out_rows.append(torch.cat(out_row, dim=3))
        return torch.cat(out_rows, dim=2)

    def forward(self, z, tiling: bool = True):
        tile_latent = self.cfg.sample_size // 8
        if tiling and (z.shape[-1] > tile_latent or z.shape[-2] > tile_latent):
            return self.tiled_decode(z)
        return self.decode(z)


def _blend_v(a, b, extent):
    # in place on b, like diffusers' blend_v: later tiles see the blended neighbour
    extent = min(a.shape[2], b.shape[2], extent)
    for y in range(extent):
        b[:, :, y, :] = a[:, :, -extent + y, :] * (1 - y / extent) + b[:, :, y, :] * (y / extent)
    return b


def _blend_h(a, b, extent):
    extent = min(a.shape[3], b.shape[3], extent)
    for x in range(extent):
        b[:, :, :, x] = a[:, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, x] * (x / extent)
    return b
