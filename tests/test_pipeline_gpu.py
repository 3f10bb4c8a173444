"""End-to-end parity of the CUDA hot path against the fp32 oracle (SURVEY.md §8c protocol):
same seed, latents, prompt embeddings and random-init weights of the named architecture;
per-step noise_pred max-rel-err <= 2e-2 (bf16), final image PSNR >= 35 dB, timesteps bit-exact."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NOISE_PRED_TOL = 2e-2     # BASELINE.json north_star: max relative error, bf16 mode
PSNR_MIN_DB = 35.0


def guided_tol(gs: float) -> float:
    """Bar for the classifier-free-GUIDED tensor eps_u + gs (eps_t - eps_u) ONLY (never for what the
    UNet outputs).  Derivation: the guided tensor is the linear map (1 - gs) eps_u + gs eps_t of the
    two UNet outputs, each of which is held to max|d| <= 2e-2 max|eps|; the worst case of the map is
    (|1 - gs| + |gs|) * 2e-2 * max|eps| = 0.28 max|eps| at gs 7.5, while with random-init text
    conditioning eps_t ~ eps_u, so max|guided| ~ max|eps|.  The bound is capped at 0.10 (measured:
    5e-2 at gs 7.5): the guided tensor is reported for information, parity is asserted on the raw
    UNet outputs — at every step on identical inputs by the teacher-forced tests below."""
    return min(NOISE_PRED_TOL * (abs(1.0 - gs) + abs(gs)), 0.10) if gs > 1.0 else NOISE_PRED_TOL


def assert_unet_outputs(raw_errs, gs):
    """FREE-RUNNING pipeline runs: only step 0 sees inputs bit-identical to the oracle's, so only step 0
    is a per-forward parity number (2e-2 bar).  Later free-running steps compare two trajectories
    (the CUDA path starts them from its own bf16-rounded latents; guidance amplifies that drift), so
    here they are only required to stay finite and bounded by the guided bar; the per-step bar is
    asserted by `teacher_forced_errors` on the oracle's own step inputs."""
    assert raw_errs[0] <= NOISE_PRED_TOL, raw_errs
    assert max(raw_errs) <= guided_tol(gs), raw_errs


def teacher_inputs(lat0, latents_after_each_step):
    """[x_0, x_1, ..., x_{n-1}]: the latents the ORACLE fed to its UNet at each step (x_0 = initial latents
    * init_noise_sigma (1.0), x_i = its scheduler output of step i-1) -> fp32 [n, B, 4, h, w]."""
    xs = [torch.as_tensor(lat0).float()] + [torch.as_tensor(x).float() for x in latents_after_each_step[:-1]]
    return torch.stack(xs)


def max_rel_err(a, b):
    """max |a-b| / max |b| (the tensor-level relative error the tolerance is quoted in)."""
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def psnr_u8(a, b):
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return 99.0 if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)


def _build(unet_cfg, vae_cfg):
    from oracle.pipeline import build_random_init
    from dreamlab_b200.engine import LCMPipelineB200
    unet, vae = build_random_init(unet_cfg, vae_cfg, seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    return unet, vae, pipe


def _run_case(unet_cfg, vae_cfg, batch, size, steps):
    from oracle.pipeline import run_pipeline, synthetic_inputs
    unet, vae, pipe = _build(unet_cfg, vae_cfg)
    pe, lat, noise = synthetic_inputs(batch, size, size, steps, ctx_dim=unet.cfg.cross_attention_dim)
    rec_o, rec_c = {}, {}
    ref_img = run_pipeline(unet, vae, pe, lat, noise, steps, 1.0, record=rec_o, tiling=False)
    img, final = pipe.generate(pe, lat, noise, steps, 1.0, record=rec_c, return_latents=True)
    torch.cuda.synchronize()
    from dreamlab_b200.scheduler import LCMSchedule
    assert LCMSchedule(steps).timesteps == rec_o["timesteps"].tolist()
    # step 0 sees identical inputs: this is the per-forward parity number
    errs = [max_rel_err(c.cpu(), o) for c, o in zip(rec_c["noise_pred"], rec_o["noise_pred"])]
    p = psnr_u8(img.cpu().numpy(), ref_img)
    print(f"noise_pred max-rel-err per step: {['%.2e' % e for e in errs]}  image PSNR {p:.1f} dB")
    return errs, p


def test_tiny_pipeline_parity():
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    errs, p = _run_case(UNetConfig.tiny(), VAEConfig.tiny(), batch=2, size=128, steps=4)
    assert max(errs) <= NOISE_PRED_TOL, errs
    assert p >= PSNR_MIN_DB, p


def test_unet_single_forward_full_arch_256():
    """Full SD1.5-LCM UNet, 256x256 (latent 32^2) so the CPU oracle finishes in seconds."""
    from oracle.unet import UNetConfig
    from oracle.pipeline import build_random_init, synthetic_inputs
    from oracle.scheduler import guidance_scale_embedding
    from dreamlab_b200.engine import UNetB200
    from dreamlab_b200.scheduler import LCMSchedule
    unet, _ = build_random_init(UNetConfig(), None, seed=0)
    eng = UNetB200(unet.state_dict(), unet.cfg, "cuda:0")
    pe, lat, _ = synthetic_inputs(1, 256, 256, 4)
    w = guidance_scale_embedding(torch.zeros(1), 256)
    with torch.no_grad():
        ref = unet(lat, torch.tensor(999), pe, w)
    kvs = eng.encode_context(pe)
    temb = eng.time_embeddings([999], 1, w.cuda())[0]
    x = lat.permute(0, 2, 3, 1).contiguous().cuda()
    eps = eng.forward(x, temb, kvs).permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    e = max_rel_err(eps.cpu(), ref)
    print(f"full-arch UNet 256^2 noise_pred max-rel-err {e:.3e}")
    assert e <= NOISE_PRED_TOL, e


def test_sd15_lcm_512_4step_vs_committed_golden():
    """BASELINE config C1 (full SD1.5-LCM arch, 512x512, 4 steps, gs 1.0, B=1) against the
    oracle outputs committed in tests/golden/ (generated by tests/golden/make_golden.py)."""
    from oracle.pipeline import build_random_init, synthetic_inputs
    from dreamlab_b200.engine import LCMPipelineB200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sd15_lcm_512_4step.npz"))
    unet, vae = build_random_init(seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    pe, lat, noise = synthetic_inputs(1, 512, 512, 4)
    rec = {}
    img = pipe.generate(pe, lat, noise, 4, 1.0, record=rec)
    torch.cuda.synchronize()
    errs = [max_rel_err(rec["noise_pred"][i].cpu(), torch.from_numpy(g["noise_pred"][i])) for i in range(4)]
    p = psnr_u8(img.cpu().numpy(), g["image"])
    print(f"C1 noise_pred max-rel-err per step: {['%.2e' % e for e in errs]}  image PSNR {p:.1f} dB")
    assert max(errs) <= NOISE_PRED_TOL, errs
    assert p >= PSNR_MIN_DB, p
    # CUDA-graph replay of the same request is byte-identical to the eager pass
    img2 = pipe.generate(pe, lat, noise, 4, 1.0, use_graph=True).clone()
    img3 = pipe.generate(pe, lat, noise, 4, 1.0, use_graph=True).clone()
    torch.cuda.synchronize()
    assert torch.equal(img2, img) and torch.equal(img3, img2)


def test_batch_invariance():
    """A request gives the same image alone or inside a batch (per-sample streams, App. A.5)."""
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from oracle.pipeline import synthetic_inputs
    _, _, pipe = _build(UNetConfig.tiny(), VAEConfig.tiny())
    pe, lat, noise = synthetic_inputs(5, 128, 128, 3, ctx_dim=64)
    full = pipe.generate(pe, lat, noise, 3, 1.0).clone()
    one = pipe.generate(pe[2:3], lat[2:3], noise[:, 2:3], 3, 1.0)
    torch.cuda.synchronize()
    d = (full[2].int() - one[0].int()).abs()
    assert d.max().item() <= 1, d.max().item()


def test_graph_batch_buckets_pad_and_share_one_pool():
    """Micro-batching produces every batch size in 1..16: graphs are captured per bucket (1/2/4/8/16), a batch
    is padded up to its bucket and the first B results returned — equal to the eager pass of the same B requests —
    and all graphs of a pipeline share one memory pool."""
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from oracle.pipeline import synthetic_inputs
    _, _, pipe = _build(UNetConfig.tiny(), VAEConfig.tiny())
    pe, lat, noise = synthetic_inputs(7, 128, 128, 3, ctx_dim=64)
    for B in (3, 5, 7, 2):
        eager = pipe.generate(pe[:B], lat[:B], noise[:, :B], 3, 1.0).clone()
        graph = pipe.generate(pe[:B], lat[:B], noise[:, :B], 3, 1.0, use_graph=True).clone()
        torch.cuda.synchronize()
        assert graph.shape == eager.shape
        assert (graph.int() - eager.int()).abs().max().item() <= 1, B          # the batch-invariance bar
    assert sorted(k[0] for k in pipe._graphs) == [2, 4, 8]                      # 3 -> 4, 5 and 7 -> 8, 2 -> 2
    assert pipe._graph_pool() is pipe._graph_pool()


def test_sd15_lcm_768_8step_vs_committed_golden():
    """BASELINE config C3 geometry (768x768, 8 steps; B=1 here): larger attention sequences
    (S = 9216 / 2304 / 576 / 144, partial tiles everywhere) and the 768^2 VAE decode (untiled:
    the variant of SURVEY.md App. A.4 whose VAE sample_size covers the request)."""
    from oracle.pipeline import build_random_init, synthetic_inputs
    from dreamlab_b200.engine import LCMPipelineB200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sd15_lcm_768_8step.npz"))
    unet, vae = build_random_init(seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    pe, lat, noise = synthetic_inputs(1, 768, 768, 8)
    rec = {}
    img = pipe.generate(pe, lat, noise, 8, 1.0, record=rec)
    torch.cuda.synchronize()
    errs = [max_rel_err(rec["noise_pred"][i].cpu(), torch.from_numpy(g["noise_pred"][i].astype(np.float32)))
            for i in range(8)]
    p = psnr_u8(img.cpu().numpy(), g["image"])
    print(f"C3 geometry noise_pred max-rel-err per step: {['%.2e' % e for e in errs]}  image PSNR {p:.1f} dB")
    assert max(errs) <= NOISE_PRED_TOL, errs
    assert p >= PSNR_MIN_DB, p


@pytest.mark.parametrize("h,w", [(32, 32), (20, 32), (32, 20), (24, 24), (28, 40)])
def test_tiled_vae_decode_matches_oracle(h, w):
    """`vae.enable_tiling()` path (reference `backends/cuda_worker.py:91`):
    latents beyond sample_size/8 decode tile by tile with the sequential in-place blend of
    diffusers' tiled_decode; ragged last tiles included (token counts that are not a multiple of
    64 — 12x12, 4x4 corner tiles — take the CUDA-core attention kernel in the mid block).  Tiny VAE: sample_size 128 -> 16-latent
    tiles at stride 12, 32-pixel blends, 96-pixel crops."""
    from oracle.pipeline import build_random_init, denormalize_to_u8
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from dreamlab_b200.engine import VAEDecoderB200
    _, vae = build_random_init(UNetConfig.tiny(), VAEConfig.tiny(), seed=0)
    dec = VAEDecoderB200(vae.state_dict(), vae.cfg, "cuda:0")
    lat = torch.randn(2, 4, h, w, generator=torch.Generator().manual_seed(5)) * 0.18215 * 3
    with torch.no_grad():
        ref = denormalize_to_u8(vae(lat / vae.cfg.scaling_factor, tiling=True))
        ref_untiled = denormalize_to_u8(vae(lat / vae.cfg.scaling_factor, tiling=False))
    x = lat.permute(0, 2, 3, 1).contiguous().cuda()
    img = dec.decode(x, tiling=True)
    img_untiled = dec.decode(x, tiling=False)
    torch.cuda.synchronize()
    assert img.shape == (2, 8 * h, 8 * w, 3)
    p, pu = psnr_u8(img.cpu().numpy(), ref), psnr_u8(img_untiled.cpu().numpy(), ref_untiled)
    cross = psnr_u8(img.cpu().numpy(), ref_untiled)
    print(f"tiled VAE {h}x{w}: PSNR tiled-vs-tiled {p:.1f} dB, untiled-vs-untiled {pu:.1f} dB, "
          f"tiled-vs-untiled-oracle {cross:.1f} dB")
    assert p >= PSNR_MIN_DB and pu >= PSNR_MIN_DB
    assert p > cross + 3.0          # it really follows the tiled arithmetic, not the untiled one


# ------------------------------------------------------------------------------------------------
# teacher forcing: every step's UNet output on the ORACLE's step inputs (north_star: "per-step
# noise_pred must match to max relative error <= 2e-2 in bf16")
# ------------------------------------------------------------------------------------------------
def test_sd15_lcm_512_teacher_forced_every_step_vs_committed_golden():
    """Config C1: the CUDA UNet is fed the oracle's latents of every step (committed golden), so each of the
    four noise_pred tensors is compared on identical inputs — no trajectory drift in the number."""
    from oracle.pipeline import build_random_init, synthetic_inputs
    from dreamlab_b200.engine import LCMPipelineB200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sd15_lcm_512_4step.npz"))
    unet, vae = build_random_init(seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    pe, lat, noise = synthetic_inputs(1, 512, 512, 4)
    rec = {}
    pipe.generate(pe, lat, noise, 4, 1.0, record=rec, teacher_latents=teacher_inputs(lat, g["latents"]))
    torch.cuda.synchronize()
    errs = [max_rel_err(rec["noise_pred"][i].cpu(), torch.from_numpy(g["noise_pred"][i])) for i in range(4)]
    lerr = [max_rel_err(rec["latents"][i].cpu(), torch.from_numpy(g["latents"][i])) for i in range(4)]
    print(f"C1 teacher-forced noise_pred max-rel-err per step: {['%.2e' % e for e in errs]}  "
          f"scheduler output {['%.2e' % e for e in lerr]}")
    assert max(errs) <= NOISE_PRED_TOL, errs
    assert max(lerr) <= NOISE_PRED_TOL, lerr
