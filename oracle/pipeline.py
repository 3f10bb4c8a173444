"""Oracle pipeline loop (test infrastructure).

Restates what `self.pipe(prompt=…, num_inference_steps=…, guidance_scale=…,
generator=gen)` does at `backends/cuda_worker.py:221-229`, following the order
of operations of the reference's in-tree copy of the loop,
`backends/rknnlcm.py:523-647` (timesteps `:559-560`, latents `:562-570`,
guidance embedding `:574-577`, loop `:586-604`, `/= scaling_factor` `:614`,
decode `:618`, postprocess `:212-264`).  fp32 on CPU.

The prompt embeddings are an *input* (north_star fixes them), and so are the
initial latents and the per-step noise draws, so that the CUDA path and the
oracle consume bit-identical randomness.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .scheduler import OracleLCMScheduler, guidance_scale_embedding
from .unet import OracleUNet, UNetConfig
from .vae import OracleVAEDecoder, VAEConfig


def build_random_init(unet_cfg: UNetConfig = None, vae_cfg: VAEConfig = None, seed: int = 0):
    """Random-init weights of the named architecture (default PyTorch inits,
    fixed seed) — the parity protocol of BASELINE.json (no checkpoints offline)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    unet = OracleUNet(unet_cfg or UNetConfig()).eval()
    vae = OracleVAEDecoder(vae_cfg or VAEConfig()).eval()
    torch.random.set_rng_state(g)
    for p in list(unet.parameters()) + list(vae.parameters()):
        p.requires_grad_(False)
    return unet, vae


def synthetic_inputs(batch: int, height: int, width: int, steps: int, ctx_dim: int = 768,
                     seed_base: int = 1000, embed_seed: int = 1):
    """prompt embeddings N(0,1) (seed 1); per-sample latents and
    per-step noise from per-sample generators seeded `seed_base+i`, drawn in the
    reference's order (latents first, then one draw per non-final step, App. A.5)."""
    g = torch.Generator().manual_seed(embed_seed)
    prompt_embeds = torch.randn(batch, 77, ctx_dim, generator=g)
    lat, noise = [], []
    for i in range(batch):
        gi = torch.Generator().manual_seed(seed_base + i)
        lat.append(torch.randn(1, 4, height // 8, width // 8, generator=gi))
        noise.append(torch.stack([torch.randn(1, 4, height // 8, width // 8, generator=gi)
                                  for _ in range(steps - 1)]) if steps > 1 else
                     torch.zeros(0, 1, 4, height // 8, width // 8))
    latents = torch.cat(lat, 0)
    step_noise = torch.cat(noise, 1) if steps > 1 else noise[0].repeat(1, batch, 1, 1, 1)
    return prompt_embeds, latents, step_noise   # [B,77,D], [B,4,h,w], [steps-1,B,4,h,w]


def denormalize_to_u8(image: torch.Tensor) -> np.ndarray:
    """`VaeImageProcessor.postprocess(output_type='pil')` up to the PIL wrap:
    clip(x/2+0.5,0,1) → NHWC → (x*255).round().astype(uint8)  (`rknnlcm.py:220-235`)."""
    x = (image / 2 + 0.5).clamp(0, 1)
    x = x.permute(0, 2, 3, 1).float().numpy()
    return (x * 255).round().astype("uint8")


@torch.no_grad()
def run_pipeline(unet: OracleUNet, vae: OracleVAEDecoder, prompt_embeds: torch.Tensor,
                 latents: torch.Tensor, step_noise: torch.Tensor, num_inference_steps: int,
                 guidance_scale: float = 1.0, output_type: str = "u8",
                 tiling: bool = True, record: Optional[dict] = None):
    """Returns uint8 NHWC images (or final latents for output_type='latent').
    `record`, if given, receives per-step noise_pred / latents for parity tests."""
    sched = OracleLCMScheduler()
    timesteps = sched.set_timesteps(num_inference_steps)
    B = latents.shape[0]
    latents = latents * sched.init_noise_sigma
    w_emb = None
    if unet.cfg.time_cond_proj_dim:
        w = torch.full((B,), guidance_scale - 1.0)
        w_emb = guidance_scale_embedding(w, unet.cfg.time_cond_proj_dim)
    denoised = latents
    if record is not None:
        record.update(timesteps=timesteps.clone(), noise_pred=[], latents=[], denoised=[])
    for i, t in enumerate(timesteps):
        eps = unet(latents, t, prompt_embeds, w_emb)
        z = step_noise[i] if i < num_inference_steps - 1 else None
        latents, denoised = sched.step(eps, int(t), latents, noise=z)
        if record is not None:
            record["noise_pred"].append(eps.clone())
            record["latents"].append(latents.clone())
            record["denoised"].append(denoised.clone())
    # diffusers' StableDiffusionPipeline decodes `latents` (== denoised after the last step)
    if output_type == "latent":
        return latents
    image = vae(latents / vae.cfg.scaling_factor, tiling=tiling)
    if record is not None:
        record["image_f32"] = image.clone()
    return denormalize_to_u8(image)


def pooled_latent_bytes(latents: torch.Tensor) -> bytes:
    """`run_job_with_latents` tail (`backends/cuda_worker.py:297-304`):
    fp32 adaptive_avg_pool2d → (8,8) → fp16 → C-order bytes (512 B for [1,4,8,8])."""
    lat8 = torch.nn.functional.adaptive_avg_pool2d(latents.to(torch.float32), (8, 8))
    return lat8.to(torch.float16).contiguous().numpy().tobytes(order="C")


def sdxl_time_ids(height: int, width: int, batch: int) -> torch.Tensor:
    """`_get_add_time_ids(original_size, crops_coords_top_left, target_size)` with the
    pipeline defaults original_size = target_size = (height, width), crop (0, 0)."""
    return torch.tensor([[height, width, 0, 0, height, width]], dtype=torch.float32).repeat(batch, 1)


@torch.no_grad()
def run_pipeline_sdxl(unet: OracleUNet, vae: OracleVAEDecoder, prompt_embeds: torch.Tensor,
                      pooled_embeds: torch.Tensor, latents: torch.Tensor, step_noise: torch.Tensor,
                      num_inference_steps: int, guidance_scale: float, height: int, width: int,
                      negative_prompt_embeds: Optional[torch.Tensor] = None,
                      negative_pooled_embeds: Optional[torch.Tensor] = None,
                      output_type: str = "u8", record: Optional[dict] = None):
    """What `self.pipe(prompt, width, height, num_inference_steps, guidance_scale, generator)`
    does at `backends/cuda_worker.py:532-539` (`StableDiffusionXLPipeline.__call__` with the
    LCMScheduler swapped in at `:387`): classifier-free guidance on a doubled batch
    [uncond, cond] when guidance_scale > 1 (SDXL-base has no time_cond_proj), text_time
    micro-conditioning, `noise_pred = uncond + gs * (text - uncond)`, `LCMScheduler.step`.
    With no negative prompt and `force_zeros_for_empty_prompt` (SDXL-base config) the
    unconditional embeddings are zeros."""
    sched = OracleLCMScheduler()
    timesteps = sched.set_timesteps(num_inference_steps)
    B = latents.shape[0]
    do_cfg = guidance_scale > 1.0 and not unet.cfg.time_cond_proj_dim
    tid = sdxl_time_ids(height, width, B)
    ctx, pooled, tids = prompt_embeds, pooled_embeds, tid
    if do_cfg:
        neg = torch.zeros_like(prompt_embeds) if negative_prompt_embeds is None else negative_prompt_embeds
        negp = torch.zeros_like(pooled_embeds) if negative_pooled_embeds is None else negative_pooled_embeds
        ctx = torch.cat([neg, prompt_embeds], 0)
        pooled = torch.cat([negp, pooled_embeds], 0)
        tids = torch.cat([tid, tid], 0)
    w_emb = None
    if unet.cfg.time_cond_proj_dim:
        w_emb = guidance_scale_embedding(torch.full((B,), guidance_scale - 1.0), unet.cfg.time_cond_proj_dim)
    latents = latents * sched.init_noise_sigma
    if record is not None:
        record.update(timesteps=timesteps.clone(), noise_pred=[], noise_pred_raw=[], latents=[])
    for i, t in enumerate(timesteps):
        x = torch.cat([latents] * 2, 0) if do_cfg else latents
        eps = unet(x, t, ctx, w_emb, text_embeds=pooled, time_ids=tids)
        if record is not None:
            record["noise_pred_raw"].append(eps.clone())       # the UNet output, before guidance
        if do_cfg:
            e_u, e_t = eps.chunk(2)
            eps = e_u + guidance_scale * (e_t - e_u)
        z = step_noise[i] if i < num_inference_steps - 1 else None
        latents, _ = sched.step(eps, int(t), latents, noise=z)
        if record is not None:
            record["noise_pred"].append(eps.clone())
            record["latents"].append(latents.clone())
    if output_type == "latent":
        return latents
    image = vae(latents / vae.cfg.scaling_factor, tiling=False)
    if record is not None:
        record["image_f32"] = image.clone()
    return denormalize_to_u8(image)
