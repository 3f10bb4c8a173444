"""BASELINE config C5 on N GPUs: SDXL-base arch (random-init), 1024x1024, CFG 7.5, LCM scheduler,
ONE image sharded over the ranks (CFG halves x row strips, NCCL over NVLink).
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_sdxl_pp.py \
      [--size 1024] [--steps 30] [--iters 3] [--no-graph] [--check]
Rank 0 prints one JSON line: seconds per image (denoise loop, max over ranks, CUDA events), the
same loop un-sharded on rank 0 for the speed-up, and (--check) the max-rel-err of the sharded
noise predictions against the un-sharded engine on the same inputs."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def log(msg):
    print(f"[rank {os.environ.get('RANK', 0)} +{time.time() - T0:6.1f}s] {msg}", file=sys.stderr, flush=True)


T0 = time.time()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--gs", type=float, default=7.5)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--tiny", action="store_true", help="tiny SDXL topology (debug)")
    ap.add_argument("--peer", action="store_true", help="exchanges as dl_peer_allgather kernels (NVLink peer memory)")
    ap.add_argument("--both", action="store_true", help="time NCCL and peer-memory exchanges in one process")
    ap.add_argument("--vae-strips", action="store_true",
                    help="also time the VAE decode as row strips over all ranks vs on one GPU, and compare the images")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        log("process group up")
        dist.barrier()
        log("first barrier done")
    from dreamlab_b200 import synthetic as syn
    from dreamlab_b200.engine import LCMPipelineB200
    from dreamlab_b200 import patch_parallel as pp
    if a.tiny:
        from oracle.unet import UNetConfig
        from oracle.vae import VAEConfig
        ucfg, vcfg = UNetConfig.tiny_sdxl(), VAEConfig.tiny()
    else:
        ucfg, vcfg = syn.sdxl_unet_cfg(), syn.sdxl_vae_cfg()
    t0 = time.time()
    pipe = LCMPipelineB200(syn.random_state_dict(syn.unet_shapes(ucfg), 0, torch.bfloat16), ucfg,
                           syn.random_state_dict(syn.vae_decoder_shapes(vcfg), 1, torch.bfloat16), vcfg, dev)
    load_s = time.time() - t0
    log(f"weights packed in {load_s:.1f}s")
    pdim = ucfg.projection_class_embeddings_input_dim - 6 * ucfg.addition_time_embed_dim
    pe, lat, noise = syn.synthetic_inputs(1, a.size, a.size, a.steps, ctx_dim=ucfg.cross_attention_dim)
    pooled = torch.randn(1, pdim, generator=torch.Generator().manual_seed(2))
    pe, lat, noise, pooled = pe.to(dev), lat.to(dev), noise.to(dev), pooled.to(dev)
    den = pp.dist_denoiser(pipe, peer=a.peer) if world > 1 else pp.PatchParallelDenoiser(pipe, pp.SingleComm())
    out = {"config": f"SDXL-base arch (random-init) {a.size}x{a.size}, {a.steps} LCM steps, CFG {a.gs}, "
                     f"1 image over {world} GPU(s)", "n_gpus": world, "load_s": round(load_s, 1),
           "exchange": "dl_peer_allgather (NVLink peer memory)" if a.peer else "ncclAllGather"}

    def timed(fn, iters):
        ts = []
        for _ in range(iters):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t))
        return sorted(ts)[len(ts) // 2]

    if a.check:
        # (1) free-running: sharded vs un-sharded trajectories (guidance amplifies bf16 drift from step 1 on);
        # (2) teacher-forced: the sharded UNet is fed the un-sharded engine's own step inputs, so every
        #     step's raw output (both CFG halves) is compared on identical inputs — the 2e-2 bar applies here.
        # Every rank runs the un-sharded engine itself (deterministic kernels: bit-identical on all GPUs).
        n_chk = min(a.steps, 3)
        rec, rec1, rec_t = {}, {}, {}
        nz = noise[:max(n_chk - 1, 1)]
        pipe.generate(pe, lat, nz, n_chk, a.gs, record=rec1, pooled_embeds=pooled)
        torch.cuda.synchronize()
        teacher = torch.stack([lat.float()] + [x.float() for x in rec1["latents"][:-1]])
        den.denoise(pe, pooled, lat, nz, n_chk, a.gs, record=rec)
        den.denoise(pe, pooled, lat, nz, n_chk, a.gs, record=rec_t, teacher_latents=teacher)
        torch.cuda.synchronize()
        log("eager sharded check passes done")
        if rank == 0:
            key = "noise_pred_raw" if a.gs > 1 else "noise_pred"
            rel = lambda x, y: float((x - y).abs().max() / y.abs().max())      # noqa: E731
            out["check_max_rel_err_vs_unsharded"] = [rel(x, y) for x, y in zip(rec[key], rec1[key])]
            out["check_teacher_forced_max_rel_err_vs_unsharded"] = [rel(x, y) for x, y in zip(rec_t[key], rec1[key])]
    if a.both and world > 1:
        # second denoiser on the same weights with the peer-memory exchanges: parity + timing
        den_p = pp.dist_denoiser(pipe, peer=True)
        rec_p = {}
        den_p.denoise(pe, pooled, lat, noise, min(a.steps, 2), a.gs, record=rec_p)
        rec_n = {}
        den.denoise(pe, pooled, lat, noise, min(a.steps, 2), a.gs, record=rec_n)
        torch.cuda.synchronize()
        out["peer_vs_nccl_bit_identical"] = all(torch.equal(x, y) for x, y in zip(rec_p["noise_pred"], rec_n["noise_pred"]))
        den_p.denoise(pe, pooled, lat, noise, a.steps, a.gs, use_graph=not a.no_graph)
        torch.cuda.synchronize()
        log("peer capture done")
    use_graph = not a.no_graph
    log(f"warm-up / capture (graph={use_graph})")
    try:
        den.denoise(pe, pooled, lat, noise, a.steps, a.gs, use_graph=use_graph)      # warm-up / capture
        torch.cuda.synchronize()
    except Exception as e:      # noqa: BLE001
        if not use_graph:
            raise
        out["graph_error"] = repr(e)[:300]
        use_graph = False
        den.denoise(pe, pooled, lat, noise, a.steps, a.gs)
        torch.cuda.synchronize()
    log("capture done, timing")
    ms = timed(lambda: den.denoise(pe, pooled, lat, noise, a.steps, a.gs, use_graph=use_graph), a.iters)
    out.update(ms_per_image_denoise=round(ms, 2), cuda_graph=use_graph,
               ms_per_unet_step=round(ms / a.steps, 3))
    if world > 1:
        ms_e = timed(lambda: den.denoise(pe, pooled, lat, noise, a.steps, a.gs, use_graph=False), 1)
        out["ms_per_image_denoise_eager"] = round(ms_e, 2)
    if a.both and world > 1:
        ms_p = timed(lambda: den_p.denoise(pe, pooled, lat, noise, a.steps, a.gs, use_graph=use_graph), a.iters)
        out["ms_per_image_denoise_peer_exchange"] = round(ms_p, 2)
        out["ms_per_unet_step_peer_exchange"] = round(ms_p / a.steps, 3)
    log("sharded timing done")
    if a.vae_strips:
        # VAE decode of the final latent: one GPU vs row strips over the world
        latf = den.denoise(pe, pooled, lat, noise, a.steps, a.gs, use_graph=use_graph).clone()
        one = pipe.vae.decode(latf)
        ms_1 = timed(lambda: pipe.vae.decode(latf), a.iters)
        out["ms_vae_decode_1gpu"] = round(ms_1, 2)
        if world > 1:
            sh = pipe.vae.decode_strips(latf, den.world)
            torch.cuda.synchronize()
            diff = (sh.int() - one.int()).abs()
            out["vae_strips_max_abs_diff_u8"] = int(diff.max())
            out["vae_strips_frac_bytes_differing"] = float((diff > 0).float().mean())
            out["ms_vae_decode_strips"] = round(timed(lambda: pipe.vae.decode_strips(latf, den.world), a.iters), 2)
    if rank == 0:
        # un-sharded reference timing of the same loop on one GPU (graph replay)
        g1 = lambda: pipe.generate(pe, lat, noise, a.steps, a.gs, pooled_embeds=pooled, use_graph=True)  # noqa: E731
        g1(); torch.cuda.synchronize()
        ts = []
        for _ in range(max(a.iters, 1)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g1(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["ms_per_image_1gpu_incl_vae"] = round(sorted(ts)[len(ts) // 2], 2)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    # destroy_process_group() hangs after NCCL collectives were captured into CUDA graphs
    # (communicator teardown waits on the captured work); the process is done, so leave.
    os._exit(0)


if __name__ == "__main__":
    main()
