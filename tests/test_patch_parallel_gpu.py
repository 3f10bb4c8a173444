"""SDXL patch-parallel denoising (BASELINE config C5 mechanics) on ONE GPU: the
ranks are threads sharing the device (`ThreadComm`), so the strip arithmetic — halo-padded convs,
split GroupNorm with gathered statistics, self-attention over gathered K/V, CFG halves on
different ranks, per-step assembly — runs through the real kernels in the ordinary `-m gpu` run.
Results must match the un-sharded engine closely and the fp32 oracle within the usual bar.
(The NCCL transport of the same code path is covered by tools/run_sdxl_pp.py on >= 2 GPUs.)"""
import threading

import pytest
import torch

from test_pipeline_gpu import NOISE_PRED_TOL, assert_unet_outputs, guided_tol, max_rel_err, teacher_inputs

pytestmark = pytest.mark.gpu

VS_UNSHARDED_TOL = NOISE_PRED_TOL   # bf16 vs bf16: GroupNorm summation order / key order differ, then bf16 rounding


def _setup(steps=2, size=128, B=1):
    from oracle.pipeline import build_random_init, synthetic_inputs
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from dreamlab_b200.engine import LCMPipelineB200
    ucfg = UNetConfig.tiny_sdxl()
    unet, vae = build_random_init(ucfg, VAEConfig.tiny(), seed=0)
    pipe = LCMPipelineB200(unet.state_dict(), unet.cfg, vae.state_dict(), vae.cfg, "cuda:0")
    pe, lat, noise = synthetic_inputs(B, size, size, steps, ctx_dim=ucfg.cross_attention_dim)
    pooled = torch.randn(B, 80, generator=torch.Generator().manual_seed(2))
    return unet, vae, pipe, (pe, pooled, lat, noise)


def _run_world(pipe, inputs, world, steps, gs, teacher=None):
    """-> per-rank (record, final latents) of a `world`-thread patch-parallel run."""
    from dreamlab_b200.patch_parallel import PatchParallelDenoiser, ThreadComm, Topology, _ThreadHub
    pe, pooled, lat, noise = inputs
    comms = ThreadComm.make(world)
    topo0 = Topology(world, 0, pipe.cfg_scale_for(gs) is not None)
    hubs = [_ThreadHub(topo0.strips) for _ in range(topo0.cfg_ways)]
    out, errs = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(0)
            den = PatchParallelDenoiser(pipe, comms[rank],
                                        lambda topo: ThreadComm(hubs[topo.cfg_index], topo.strip_index))
            rec = {}
            final = den.denoise(pe, pooled, lat, noise, steps, gs, record=rec, teacher_latents=teacher)
            torch.cuda.synchronize()
            out[rank] = (rec, final)
        except BaseException as e:        # noqa: BLE001
            import traceback
            errs.append((rank, traceback.format_exc()))
            for c in comms:
                c.hub.barrier.abort()
            for h in hubs:
                h.barrier.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(300) for t in ts]
    assert not errs, errs[0][1]
    return out


@pytest.mark.parametrize("world,gs", [(1, 7.5), (2, 1.0), (2, 7.5), (4, 7.5), (4, 1.0)])
def test_patch_parallel_matches_unsharded_and_oracle(world, gs):
    from oracle.pipeline import run_pipeline_sdxl
    steps, size = 2, 128
    unet, vae, pipe, inputs = _setup(steps, size)
    pe, pooled, lat, noise = inputs
    rec_o, rec_1 = {}, {}
    run_pipeline_sdxl(unet, vae, pe, pooled, lat, noise, steps, gs, size, size, record=rec_o,
                      output_type="latent")
    pipe.generate(pe, lat, noise, steps, gs, record=rec_1, pooled_embeds=pooled)
    torch.cuda.synchronize()
    if world == 1:
        from dreamlab_b200.patch_parallel import PatchParallelDenoiser, SingleComm
        rec = {}
        den = PatchParallelDenoiser(pipe, SingleComm())
        den.force_split = True
        final = den.denoise(pe, pooled, lat, noise, steps, gs, record=rec)
        out = [(rec, final)]
    else:
        out = _run_world(pipe, inputs, world, steps, gs)
    for rank, (rec, final) in enumerate(out):
        key = "noise_pred_raw" if gs > 1.0 else "noise_pred"      # the UNet output (before guidance)
        e1 = [max_rel_err(a, b) for a, b in zip(rec[key], rec_1[key])]
        eo = [max_rel_err(a.cpu(), b) for a, b in zip(rec[key], rec_o["noise_pred_raw"])]
        eg = [max_rel_err(a.cpu(), b) for a, b in zip(rec["noise_pred"], rec_o["noise_pred"])]
        if rank == 0:
            print(f"world={world} gs={gs}: UNet output vs un-sharded {['%.2e' % e for e in e1]}  vs oracle "
                  f"{['%.2e' % e for e in eo]}  guided vs oracle {['%.2e' % e for e in eg]}")
        assert e1[0] <= VS_UNSHARDED_TOL and max(e1) <= guided_tol(gs), (rank, e1)
        assert_unet_outputs(eo, gs)
        assert max(eg) <= guided_tol(gs), (rank, eg)
        # every rank holds bit-identical latents (same kernels on identical gathered inputs)
        assert torch.equal(final, out[0][1])


@pytest.mark.parametrize("world,gs", [(2, 7.5), (4, 7.5), (4, 1.0)])
def test_patch_parallel_teacher_forced_every_step(world, gs):
    """Row strips + CFG halves, every step on the ORACLE's step inputs: raw UNet outputs (both CFG halves)
    against the oracle AND against the un-sharded engine (teacher-forced as well) at the 2e-2 bar — the
    free-running comparison above measures how guidance amplifies bf16 trajectory drift, this one the
    sharded forward itself."""
    from oracle.pipeline import run_pipeline_sdxl
    steps, size = 3, 128
    unet, vae, pipe, inputs = _setup(steps, size)
    pe, pooled, lat, noise = inputs
    rec_o, rec_1 = {}, {}
    run_pipeline_sdxl(unet, vae, pe, pooled, lat, noise, steps, gs, size, size, record=rec_o, output_type="latent")
    teacher = teacher_inputs(lat, rec_o["latents"])
    pipe.generate(pe, lat, noise, steps, gs, record=rec_1, pooled_embeds=pooled, teacher_latents=teacher)
    torch.cuda.synchronize()
    out = _run_world(pipe, inputs, world, steps, gs, teacher=teacher)
    key = "noise_pred_raw" if gs > 1.0 else "noise_pred"
    for rank, (rec, _) in enumerate(out):
        e1 = [max_rel_err(a, b) for a, b in zip(rec[key], rec_1[key])]
        eo = [max_rel_err(a.cpu(), b) for a, b in zip(rec[key], rec_o["noise_pred_raw"])]
        if rank == 0:
            print(f"world={world} gs={gs} teacher-forced: UNet output vs un-sharded {['%.2e' % e for e in e1]}  "
                  f"vs oracle {['%.2e' % e for e in eo]}")
        assert max(e1) <= NOISE_PRED_TOL, (rank, e1)
        assert max(eo) <= NOISE_PRED_TOL, (rank, eo)


def test_split_groupnorm_matches_fused():
    """dl_groupnorm_stats + dl_groupnorm_apply (1 rank) vs the fused cooperative kernel, incl. the
    two-source concat and a halo-padded destination."""
    from dreamlab_b200 import lib
    torch.manual_seed(0)
    for (B, hw, c0, c1, silu) in [(2, 1024, 320, 0, True), (1, 4096, 640, 320, True), (3, 256, 1280, 1280, False),
                                  (2, 16384, 128, 0, True)]:
        C = c0 + c1
        x0 = (torch.randn(B, hw, c0, device="cuda") * 2 + 0.5).bfloat16()
        x1 = (torch.randn(B, hw, c1, device="cuda") - 0.3).bfloat16() if c1 else None
        gw = torch.randn(C, device="cuda")
        gb = torch.randn(C, device="cuda")
        ref = torch.empty(B, hw, C, device="cuda", dtype=torch.bfloat16)
        ws = torch.empty(lib.groupnorm_workspace_bytes(B, 32), device="cuda", dtype=torch.uint8)
        lib.groupnorm(x0, ref, gw, gb, ws, nimg=B, hw=hw, groups=32, eps=1e-5, silu=silu, x1=x1)
        ws2 = torch.zeros(lib.groupnorm_split_workspace_bytes(B, 32), device="cuda", dtype=torch.uint8)
        stats = torch.empty(B, 32, 2, device="cuda")
        lib.groupnorm_stats(x0, stats, ws2, nimg=B, hw=hw, groups=32, x1=x1)
        W = 32
        H = hw // W
        pad = torch.zeros(B, H + 2, W, C, device="cuda", dtype=torch.bfloat16)
        lib.groupnorm_apply(x0, pad[:, 1:], gw, gb, stats.unsqueeze(0), nimg=B, hw=hw, groups=32, eps=1e-5,
                            silu=silu, x1=x1, out_img_stride=(H + 2) * W * C)
        torch.cuda.synchronize()
        got = pad[:, 1:H + 1].reshape(B, hw, C)
        assert float(pad[:, 0].abs().max()) == 0 and float(pad[:, H + 1].abs().max()) == 0
        err = (got.float() - ref.float()).abs().max().item()
        assert err <= 0.0625, (B, hw, c0, c1, err)     # <= 1 bf16 ulp at |y| < 8
        # the stats kernel left its counters clean: a second run gives the same record
        stats2 = torch.empty_like(stats)
        lib.groupnorm_stats(x0, stats2, ws2, nimg=B, hw=hw, groups=32, x1=x1)
        torch.cuda.synchronize()
        assert torch.equal(stats, stats2)


def test_halo_conv_matches_dense_rows():
    """conv3x3 over a halo-padded strip == the same rows of the dense conv: bit-exact when both
    take the same kernel path (>= 16 rows: halo-patch mode, the exchanged rows just replace TMA
    zero fill), within bf16 rounding when the short strip falls back to the per-tap loader (the
    fp32 accumulation order over taps / channel chunks differs)."""
    from dreamlab_b200 import lib
    torch.manual_seed(0)
    B, H, W, C, N = 2, 32, 32, 128, 192
    x = torch.randn(B, H, W, C, device="cuda").bfloat16()
    w = (torch.randn(N, 9 * C, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    dense = torch.empty(B, H, W, N, device="cuda", dtype=torch.bfloat16)
    lib.igemm(x, w, dense, nimg=B, h=H, w=W, taps=9, n=N, bias=bias)
    for r0, hl in [(0, 16), (16, 16), (8, 16), (4, 8), (28, 4)]:
        pad = torch.zeros(B, hl + 2, W, C, device="cuda", dtype=torch.bfloat16)
        lo, hi = max(r0 - 1, 0), min(r0 + hl + 1, H)
        pad[:, lo - (r0 - 1):hi - (r0 - 1)] = x[:, lo:hi]
        out = torch.empty(B, hl, W, N, device="cuda", dtype=torch.bfloat16)
        lib.igemm(pad, w, out, nimg=B, h=hl, w=W, taps=9, n=N, bias=bias, in_rows=hl + 2, in_row0=1)
        torch.cuda.synchronize()
        ref = dense[:, r0:r0 + hl]
        if hl >= 16:
            assert torch.equal(out, ref), (r0, hl)
        else:
            assert torch.allclose(out.float(), ref.float(), rtol=2e-2, atol=2e-2), (r0, hl)


def _decode_world(dec, x, world, f32_out=False):
    """-> per-rank result of `decode_strips` over `world` thread ranks sharing cuda:0."""
    from dreamlab_b200.patch_parallel import SingleComm, ThreadComm
    if world == 1:
        return [dec.decode_strips(x, SingleComm(), f32_out=f32_out)]
    comms = ThreadComm.make(world)
    out, errs = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(0)
            out[rank] = dec.decode_strips(x, comms[rank], f32_out=f32_out)
            torch.cuda.synchronize()
        except BaseException:             # noqa: BLE001
            import traceback
            errs.append(traceback.format_exc())
            comms[0].hub.barrier.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(300) for t in ts]
    assert not errs, errs[0]
    return out


@pytest.mark.parametrize("world,h,w,B", [(1, 16, 16, 1), (2, 16, 16, 2), (4, 16, 16, 1), (2, 32, 24, 1),
                                         (4, 32, 32, 1), (8, 32, 16, 1)])
def test_vae_strip_decode_matches_unsharded_and_oracle(world, h, w, B):
    """SURVEY.md §8e "VAE decode in C5": the decoder over row strips — halo rows in front of every
    3x3 conv (upsample convs included), split GroupNorm, mid-block attention over the gathered
    tokens, pixel-row gather — against the one-GPU decode of the same kernels (GroupNorm merge /
    key order differ, then bf16 rounding) and against the fp32 oracle at the pipeline's PSNR bar.
    (4, 16, 16): 64 query tokens per rank; (8, 32, 16): 4 latent rows x 16 = 64 tokens per rank and 512 keys."""
    from oracle.pipeline import build_random_init, denormalize_to_u8
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from dreamlab_b200.engine import VAEDecoderB200
    from test_pipeline_gpu import PSNR_MIN_DB, psnr_u8
    _, vae = build_random_init(UNetConfig.tiny(), VAEConfig.tiny(), seed=0)
    dec = VAEDecoderB200(vae.state_dict(), vae.cfg, "cuda:0")
    lat = torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(7)) * 0.18215 * 3
    with torch.no_grad():
        ref = denormalize_to_u8(vae(lat / vae.cfg.scaling_factor, tiling=False))
    x = lat.permute(0, 2, 3, 1).contiguous().cuda()
    one = dec.decode(x)
    torch.cuda.synchronize()
    outs = _decode_world(dec, x, world)
    for r, img in enumerate(outs):
        assert img.shape == (B, 8 * h, 8 * w, 3) and img.dtype == torch.uint8
        assert torch.equal(img, outs[0]), r                  # every rank holds the same gathered image
    got = outs[0].cpu().numpy()
    p_one, p_ref = psnr_u8(got, one.cpu().numpy()), psnr_u8(got, ref)
    dmax = int((outs[0].int() - one.int()).abs().max())
    print(f"VAE strips world={world} {h}x{w} B={B}: vs one-GPU decode PSNR {p_one:.1f} dB (max |d| {dmax}), "
          f"vs oracle {p_ref:.1f} dB (one-GPU vs oracle {psnr_u8(one.cpu().numpy(), ref):.1f} dB)")
    assert p_ref >= PSNR_MIN_DB, p_ref
    assert p_one >= 45.0, p_one


def test_strips_refuse_fp32_mode():
    """The split GroupNorm kernels exist for bf16 only: the fp32 precision mode (a one-GPU parity
    mode) must refuse a strip communicator loudly instead of computing garbage."""
    from oracle.pipeline import build_random_init
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from dreamlab_b200.engine import VAEDecoderB200
    from dreamlab_b200.patch_parallel import SingleComm
    _, vae = build_random_init(UNetConfig.tiny(), VAEConfig.tiny(), seed=0)
    dec = VAEDecoderB200(vae.state_dict(), vae.cfg, "cuda:0", precision="fp32")
    x = torch.zeros(1, 16, 16, 4, device="cuda")
    with pytest.raises(RuntimeError, match="bf16 path only"):
        dec.decode_strips(x, SingleComm())
