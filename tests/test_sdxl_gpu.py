"""SDXL-class path (BASELINE config C5 architecture) against the fp32 oracle:
text_time micro-conditioning, per-level head counts (head dim 64), deep transformers with linear
projections, classifier-free guidance on the doubled batch, LCMScheduler.  Same protocol and
tolerances as tests/test_pipeline_gpu.py."""
import os

import numpy as np
import pytest
import torch

from test_pipeline_gpu import (NOISE_PRED_TOL, PSNR_MIN_DB, assert_unet_outputs, guided_tol, max_rel_err, psnr_u8,
                               teacher_inputs)

pytestmark = pytest.mark.gpu


def _inputs(batch, size, steps, ctx_dim, pooled_dim):
    from oracle.pipeline import synthetic_inputs
    pe, lat, noise = synthetic_inputs(batch, size, size, steps, ctx_dim=ctx_dim)
    pooled = torch.randn(batch, pooled_dim, generator=torch.Generator().manual_seed(2))
    return pe, pooled, lat, noise


@pytest.mark.parametrize("gs", [7.5, 1.0])
def test_tiny_sdxl_pipeline_parity(gs):
    """SDXL topology at small widths, live against the oracle; gs=7.5 runs CFG (doubled batch),
    gs=1.0 the un-guided branch of the same pipeline."""
    from oracle.pipeline import build_random_init, run_pipeline_sdxl
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from dreamlab_b200.engine import LCMPipelineB200
    ucfg = UNetConfig.tiny_sdxl()
    unet, vae = build_random_init(ucfg, VAEConfig.tiny(), seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    assert pipe.is_sdxl
    steps, size, B = 3, 128, 2
    pdim = ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim
    pe, pooled, lat, noise = _inputs(B, size, steps, ucfg.cross_attention_dim, pdim)
    rec_o, rec_c = {}, {}
    ref = run_pipeline_sdxl(unet, vae, pe, pooled, lat, noise, steps, gs, size, size, record=rec_o)
    img = pipe.generate(pe, lat, noise, steps, gs, record=rec_c, pooled_embeds=pooled)
    torch.cuda.synchronize()
    errs = [max_rel_err(c.cpu(), o) for c, o in zip(rec_c["noise_pred"], rec_o["noise_pred"])]
    raw = [max_rel_err(c.cpu(), o) for c, o in zip(rec_c.get("noise_pred_raw", rec_c["noise_pred"]),
                                                   rec_o["noise_pred_raw"])]
    p = psnr_u8(img.cpu().numpy(), ref)
    print(f"tiny SDXL gs={gs}: UNet output max-rel-err per step {['%.2e' % e for e in raw]}  guided "
          f"{['%.2e' % e for e in errs]}  PSNR {p:.1f} dB")
    assert_unet_outputs(raw, gs)
    assert max(errs) <= guided_tol(gs), errs
    assert p >= PSNR_MIN_DB, p
    # graph replay == eager, byte for byte
    img2 = pipe.generate(pe, lat, noise, steps, gs, use_graph=True, pooled_embeds=pooled).clone()
    torch.cuda.synchronize()
    assert torch.equal(img2, img)


def test_sdxl_base_512_3step_cfg_vs_committed_golden():
    """Full SDXL-base UNet (2.57 B parameters, random-init seed 0) + SDXL VAE scaling, 512x512,
    3 LCM steps, CFG 7.5, against oracle outputs committed in tests/golden/ (make_golden.py --sdxl)."""
    from dreamlab_b200 import synthetic as syn
    from dreamlab_b200.engine import LCMPipelineB200
    from oracle.pipeline import build_random_init
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sdxl_512_3step_cfg.npz"))
    unet, vae = build_random_init(UNetConfig.sdxl_base(), VAEConfig(scaling_factor=0.13025, sample_size=1024), seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    del unet
    pe, pooled, lat, noise = _inputs(1, 512, 3, 2048, 1280)
    rec = {}
    img = pipe.generate(pe, lat, noise, 3, 7.5, record=rec, pooled_embeds=pooled)
    torch.cuda.synchronize()
    errs = [max_rel_err(rec["noise_pred"][i].cpu(), torch.from_numpy(g["noise_pred"][i])) for i in range(3)]
    raw = [max_rel_err(rec["noise_pred_raw"][i].cpu(), torch.from_numpy(g["noise_pred_raw"][i])) for i in range(3)]
    p = psnr_u8(img.cpu().numpy(), g["image"])
    print(f"SDXL 512^2 CFG: UNet output max-rel-err per step {['%.2e' % e for e in raw]}  guided "
          f"{['%.2e' % e for e in errs]}  image PSNR {p:.1f} dB")
    assert_unet_outputs(raw, 7.5)
    assert max(errs) <= guided_tol(7.5), errs
    assert p >= PSNR_MIN_DB, p


# ------------------------------------------------------------------------------------------------
# teacher forcing under classifier-free guidance: BOTH halves of the raw UNet output, EVERY step,
# on the oracle's own step inputs, at the 2e-2 bar (no guided_tol on what the UNet outputs)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gs", [7.5, 1.0])
def test_tiny_sdxl_teacher_forced_every_step(gs):
    from oracle.pipeline import build_random_init, run_pipeline_sdxl
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from dreamlab_b200.engine import LCMPipelineB200
    ucfg = UNetConfig.tiny_sdxl()
    unet, vae = build_random_init(ucfg, VAEConfig.tiny(), seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    steps, size, B = 4, 128, 2
    pdim = ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim
    pe, pooled, lat, noise = _inputs(B, size, steps, ucfg.cross_attention_dim, pdim)
    rec_o, rec_c = {}, {}
    run_pipeline_sdxl(unet, vae, pe, pooled, lat, noise, steps, gs, size, size, record=rec_o, output_type="latent")
    pipe.generate(pe, lat, noise, steps, gs, record=rec_c, pooled_embeds=pooled,
                  teacher_latents=teacher_inputs(lat, rec_o["latents"]))
    torch.cuda.synchronize()
    raw = [max_rel_err(c.cpu(), o) for c, o in zip(rec_c.get("noise_pred_raw", rec_c["noise_pred"]),
                                                   rec_o["noise_pred_raw"])]
    halves = []
    if gs > 1.0:                       # the unconditional and the text half separately
        for c, o in zip(rec_c["noise_pred_raw"], rec_o["noise_pred_raw"]):
            halves.append((max_rel_err(c[:B].cpu(), o[:B]), max_rel_err(c[B:].cpu(), o[B:])))
    guided = [max_rel_err(c.cpu(), o) for c, o in zip(rec_c["noise_pred"], rec_o["noise_pred"])]
    print(f"tiny SDXL gs={gs} teacher-forced: UNet output per step {['%.2e' % e for e in raw]}  halves "
          f"{[('%.2e' % a, '%.2e' % b) for a, b in halves]}  guided {['%.2e' % e for e in guided]}")
    assert max(raw) <= NOISE_PRED_TOL, raw
    assert all(max(h) <= NOISE_PRED_TOL for h in halves), halves
    assert max(guided) <= guided_tol(gs), guided


def test_sdxl_base_512_teacher_forced_every_step_vs_committed_golden():
    """Full SDXL-base UNet, CFG 7.5: the three UNet forwards on the oracle's committed step latents."""
    from dreamlab_b200.engine import LCMPipelineB200
    from oracle.pipeline import build_random_init
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sdxl_512_3step_cfg.npz"))
    unet, vae = build_random_init(UNetConfig.sdxl_base(), VAEConfig(scaling_factor=0.13025, sample_size=1024), seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    del unet
    pe, pooled, lat, noise = _inputs(1, 512, 3, 2048, 1280)
    rec = {}
    pipe.generate(pe, lat, noise, 3, 7.5, record=rec, pooled_embeds=pooled,
                  teacher_latents=teacher_inputs(lat, g["latents"]))
    torch.cuda.synchronize()
    raw = [max_rel_err(rec["noise_pred_raw"][i].cpu(), torch.from_numpy(g["noise_pred_raw"][i])) for i in range(3)]
    halves = [(max_rel_err(rec["noise_pred_raw"][i][:1].cpu(), torch.from_numpy(g["noise_pred_raw"][i][:1])),
               max_rel_err(rec["noise_pred_raw"][i][1:].cpu(), torch.from_numpy(g["noise_pred_raw"][i][1:])))
              for i in range(3)]
    guided = [max_rel_err(rec["noise_pred"][i].cpu(), torch.from_numpy(g["noise_pred"][i])) for i in range(3)]
    print(f"SDXL 512^2 CFG teacher-forced: UNet output per step {['%.2e' % e for e in raw]}  halves "
          f"{[('%.2e' % a, '%.2e' % b) for a, b in halves]}  guided {['%.2e' % e for e in guided]}")
    assert max(raw) <= NOISE_PRED_TOL, raw
    assert all(max(h) <= NOISE_PRED_TOL for h in halves), halves
    assert max(guided) <= guided_tol(7.5), guided
