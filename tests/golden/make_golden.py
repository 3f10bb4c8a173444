"""Generates the committed golden fixtures from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot produce vectors here:
its arithmetic lives in `diffusers`, which is absent from this image."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.pipeline import build_random_init, run_pipeline, synthetic_inputs  # noqa: E402
from oracle.unet import UNetConfig  # noqa: E402
from oracle.vae import VAEConfig  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tiny():
    unet, vae = build_random_init(UNetConfig.tiny(), VAEConfig.tiny(), seed=0)
    pe, lat, noise = synthetic_inputs(1, 64, 64, 2, ctx_dim=unet.cfg.cross_attention_dim)
    rec = {}
    img = run_pipeline(unet, vae, pe, lat, noise, 2, 1.0, record=rec, tiling=False)
    np.savez_compressed(os.path.join(HERE, "tiny_pipeline.npz"),
                        noise_pred0=rec["noise_pred"][0].numpy(),
                        final_latents=rec["latents"][-1].numpy(), image=img)


def full(size=512, steps=4, tiling=False):
    """BASELINE config C1: SD1.5-LCM arch, random-init seed 0, 512x512, 4 steps, gs 1.0, B=1."""
    torch.set_num_threads(os.cpu_count())
    unet, vae = build_random_init(seed=0)
    pe, lat, noise = synthetic_inputs(1, size, size, steps)
    rec = {}
    img = run_pipeline(unet, vae, pe, lat, noise, steps, 1.0, record=rec, tiling=tiling)
    # noise_pred / latents as fp16 keep the fixture small (they are compared at 2e-2)
    np.savez_compressed(os.path.join(HERE, f"sd15_lcm_{size}_{steps}step.npz"),
                        noise_pred=torch.stack(rec["noise_pred"]).numpy().astype(np.float32 if size <= 512 else np.float16),
                        latents=torch.stack(rec["latents"]).numpy().astype(np.float32 if size <= 512 else np.float16),
                        image=img)


def sdxl_inputs(batch, size, steps, ctx_dim=2048, pooled_dim=1280):
    pe, lat, noise = synthetic_inputs(batch, size, size, steps, ctx_dim=ctx_dim)
    pooled = torch.randn(batch, pooled_dim, generator=torch.Generator().manual_seed(2))
    return pe, pooled, lat, noise


def sdxl(size=512, steps=3, gs=7.5):
    """BASELINE config C5 architecture (SDXL-base UNet, random-init seed 0; SDXL VAE scaling),
    classifier-free guidance 7.5 with the zero unconditional embedding, at a size and step
    count the fp32 CPU oracle finishes in minutes."""
    from oracle.pipeline import run_pipeline_sdxl
    torch.set_num_threads(os.cpu_count())
    unet, vae = build_random_init(UNetConfig.sdxl_base(), VAEConfig(scaling_factor=0.13025, sample_size=1024), seed=0)
    pe, pooled, lat, noise = sdxl_inputs(1, size, steps)
    rec = {}
    img = run_pipeline_sdxl(unet, vae, pe, pooled, lat, noise, steps, gs, size, size, record=rec)
    np.savez_compressed(os.path.join(HERE, f"sdxl_{size}_{steps}step_cfg.npz"),
                        noise_pred=torch.stack(rec["noise_pred"]).numpy().astype(np.float32),
                        noise_pred_raw=torch.stack(rec["noise_pred_raw"]).numpy().astype(np.float32),
                        latents=torch.stack(rec["latents"]).numpy().astype(np.float32), image=img)


if __name__ == "__main__":
    if len(sys.argv) == 1 or "--tiny" in sys.argv:
        tiny()
    if "--full" in sys.argv:
        full()
    if "--sdxl" in sys.argv:
        sdxl()
    if "--c3" in sys.argv:
        full(768, 8)          # BASELINE config C3 geometry (B=1), untiled VAE decode
