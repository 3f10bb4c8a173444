"""Oracle UNet2DConditionModel — SD1.5 (+LCM time_cond_proj_dim) and SDXL-base (test infrastructure).

fp32 PyTorch restatement of the module the reference runs through diffusers at
`backends/cuda_worker.py:222` (UNet call mirrored at `backends/rknnlcm.py:588-593`).
Architecture per SURVEY.md Appendix A.2; parameter names are the diffusers
state-dict keys so a real checkpoint loads 1:1.  NCHW like diffusers.  The SDXL variant is the
UNet `StableDiffusionXLPipeline` runs at `backends/cuda_worker.py:532` (config C5).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_attn: Tuple[bool, ...] = (True, True, True, False)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: int = 8          # == number of heads (diffusers naming quirk)
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    time_cond_proj_dim: Optional[int] = 256   # LCM_Dreamshaper_v7; None for vanilla SD1.5
    # --- SDXL additions (App. A.2 "SDXL-base"); defaults reproduce SD1.5
    transformer_layers_per_block: Tuple[int, ...] = ()     # () == 1 everywhere
    use_linear_projection: bool = False
    addition_embed_type: Optional[str] = None              # "text_time" for SDXL
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816

    def heads_at(self, i: int) -> int:
        a = self.attention_head_dim
        return a if isinstance(a, int) else a[i]

    def depth_at(self, i: int) -> int:
        t = self.transformer_layers_per_block
        return t[i] if t else 1

    @staticmethod
    def sd15_lcm() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def sdxl_base() -> "UNetConfig":
        return UNetConfig(block_out_channels=(320, 640, 1280), down_attn=(False, True, True),
                          cross_attention_dim=2048, attention_head_dim=(5, 10, 20),
                          time_cond_proj_dim=None, transformer_layers_per_block=(1, 2, 10),
                          use_linear_projection=True, addition_embed_type="text_time")

    @staticmethod
    def tiny_sdxl() -> "UNetConfig":
        """SDXL topology (3 levels, no attention at level 0, deep transformers, linear
        projections, text_time embedding), small widths."""
        return UNetConfig(block_out_channels=(64, 128, 256), down_attn=(False, True, True),
                          cross_attention_dim=128, attention_head_dim=(2, 2, 4),
                          time_cond_proj_dim=None, transformer_layers_per_block=(1, 2, 3),
                          use_linear_projection=True, addition_embed_type="text_time",
                          addition_time_embed_dim=32, projection_class_embeddings_input_dim=6 * 32 + 80)

    @staticmethod
    def tiny() -> "UNetConfig":
        """Same topology, small widths: CPU-fast structural tests."""
        return UNetConfig(block_out_channels=(64, 128, 256, 256), cross_attention_dim=64,
                          attention_head_dim=4, time_cond_proj_dim=32)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0), fp32."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
    emb = t.to(torch.float32)[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim, cond_proj_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)
        self.cond_proj = nn.Linear(cond_proj_dim, in_dim, bias=False) if cond_proj_dim else None

    def forward(self, sample, cond=None):
        if cond is not None and self.cond_proj is not None:
            sample = sample + self.cond_proj(cond)
        return self.linear_2(F.silu(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout) if temb_dim else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, heads, ctx_dim=None, qkv_bias=False):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=qkv_bias)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=qkv_bias)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=qkv_bias)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, S, C = x.shape
        h = self.heads
        q = self.to_q(x).view(B, S, h, C // h).transpose(1, 2)
        k = self.to_k(ctx).view(B, -1, h, C // h).transpose(1, 2)
        v = self.to_v(ctx).view(B, -1, h, C // h).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(C // h), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, C)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        # net.0 = GEGLU, net.1 = Dropout (no params), net.2 = Linear
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, ctx_dim, groups, depth=1, linear_proj=False):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        # use_linear_projection: SD1.5 False (1x1 conv before the reshape), SDXL True (Linear after)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.linear_proj:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if self.linear_proj:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(B, H, W, C).permute(0, 3, 1, 2))
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, cin, cout, temb, attn, down, heads=None, depth=1):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, cfg.norm_num_groups, cfg.norm_eps)
             for i in range(cfg.layers_per_block)])
        if attn:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(cout, heads, cfg.cross_attention_dim, cfg.norm_num_groups, depth,
                                    cfg.use_linear_projection) for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, c, temb, heads, depth):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(c, c, temb, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(c, heads, cfg.cross_attention_dim, cfg.norm_num_groups, depth,
                                cfg.use_linear_projection)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, prev_out, cout, skip_chs, temb, attn, up, heads=None, depth=1):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList(
            [ResnetBlock2D((prev_out if i == 0 else cout) + skip_chs[i], cout, temb,
                           cfg.norm_num_groups, cfg.norm_eps) for i in range(n)])
        if attn:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(cout, heads, cfg.cross_attention_dim, cfg.norm_num_groups, depth,
                                    cfg.use_linear_projection) for _ in range(n)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class OracleUNet(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb, cfg.time_cond_proj_dim)
        if cfg.addition_embed_type == "text_time":
            # add_time_proj = Timesteps(addition_time_embed_dim) has no parameters
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, temb, None)
        else:
            self.add_embedding = None
        self.down_blocks = nn.ModuleList()
        skip_chs = [ch[0]]
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            last = i == len(ch) - 1
            self.down_blocks.append(DownBlock(cfg, cin, cout, temb, cfg.down_attn[i], not last,
                                              cfg.heads_at(i), cfg.depth_at(i)))
            skip_chs += [cout] * cfg.layers_per_block + ([cout] if not last else [])
        n_lv = len(ch)
        self.mid_block = MidBlock(cfg, ch[-1], temb, cfg.heads_at(n_lv - 1), cfg.depth_at(n_lv - 1))
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        rev_attn = list(reversed(cfg.down_attn))
        prev = rev[0]
        for i, c in enumerate(rev):
            n = cfg.layers_per_block + 1
            sk = [skip_chs.pop() for _ in range(n)]
            last = i == len(ch) - 1
            lv = n_lv - 1 - i
            self.up_blocks.append(UpBlock(cfg, prev, c, sk, temb, rev_attn[i], not last,
                                          cfg.heads_at(lv), cfg.depth_at(lv)))
            prev = c
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, timestep_cond=None,
                text_embeds=None, time_ids=None):
        B = sample.shape[0]
        t = torch.as_tensor(timestep).reshape(-1).expand(B)
        t_emb = timestep_embedding(t, self.cfg.block_out_channels[0])
        emb = self.time_embedding(t_emb, timestep_cond)
        if self.add_embedding is not None:
            # SDXL "text_time": sinusoid of each of the 6 micro-conditioning ids, concatenated
            # after the pooled text embedding (text first), through its own 2-layer MLP
            te = timestep_embedding(time_ids.reshape(-1), self.cfg.addition_time_embed_dim)
            add = torch.cat([text_embeds.to(torch.float32), te.reshape(B, -1)], dim=-1)
            emb = emb + self.add_embedding(add)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        return self.conv_out(F.silu(self.conv_norm_out(x)))
