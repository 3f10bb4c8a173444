"""The b200 worker behind the reference's PipelineWorker contract — the behaviours the reference
pins in `tests/test_sdxl_worker.py:118-298` (attributes, PNG magic, seed echo, same-seed
byte-identity, 512-byte latents, size errors, seed=None), on a random-init fixture model."""
import io
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


class Req(SimpleNamespace):
    pass


def job(prompt="a cat", size="128x128", steps=2, gs=1.0, seed=42):
    return SimpleNamespace(req=Req(prompt=prompt, size=size, num_inference_steps=steps,
                                   guidance_scale=gs, seed=seed))


@pytest.fixture(scope="module")
def worker(tmp_path_factory):
    from dreamlab_b200 import synthetic as S
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    root = tmp_path_factory.mktemp("models")
    ucfg = UNetConfig.tiny()
    ucfg.cross_attention_dim = 768            # the factory routes on 768/1024 -> sd15
    S.write_model_dir(str(root / "tiny-lcm"), ucfg, VAEConfig.tiny())
    os.environ["MODEL_ROOT"], os.environ["MODEL"] = str(root), "tiny-lcm"
    os.environ.pop("CUDA_DEVICE", None)
    from backends.worker_factory import create_cuda_worker
    return create_cuda_worker(worker_id=0)


def test_attributes(worker):
    from backends.b200_worker import B200Worker
    assert isinstance(worker, B200Worker)
    assert worker.worker_id == 0 and worker.pipe is not None
    assert str(worker.device).startswith("cuda") and worker.dtype == torch.float16


def test_png_seed_and_determinism(worker):
    from PIL import Image
    png, seed = worker.run_job(job(seed=42))
    assert png[:8] == b"\x89PNG\r\n\x1a\n" and seed == 42
    assert Image.open(io.BytesIO(png)).size == (128, 128)
    png2, _ = worker.run_job(job(seed=42))
    assert png2 == png                                   # same seed => byte-identical PNG
    png3, _ = worker.run_job(job(seed=43))
    assert png3 != png


def test_latents_bytes(worker):
    png, seed, lat = worker.run_job_with_latents(job(seed=7))
    assert len(lat) == 512 and seed == 7 and png[:4] == b"\x89PNG"
    assert worker.run_job(job(seed=7))[0] == png          # same pass, same image


def test_sizes_and_errors(worker):
    from PIL import Image
    for s in ("64x64", "128x64", "192x192"):
        png, _ = worker.run_job(job(size=s))
        w, h = (int(v) for v in s.split("x"))
        assert Image.open(io.BytesIO(png)).size == (w, h)
    with pytest.raises(RuntimeError, match="Invalid size"):
        worker.run_job(job(size="banana"))
    # sizes the UNet cannot halve exactly at every level are refused up front with a clear message
    # (they used to fail deep in the pass with 'im2col_s2: bad args')
    with pytest.raises(RuntimeError, match="multiples of"):
        worker.run_job(job(size="72x72"))
    a = worker.run_job(job(seed=None))
    b = worker.run_job(job(seed=None))
    assert a[1] != b[1] and a[0] != b[0]


def test_batch_equals_singles_and_pool_roundtrip(worker):
    from backends.worker_pool import GenerationJob, WorkerPool
    from unittest.mock import Mock
    jobs = [job(prompt=f"p{i}", seed=100 + i) for i in range(4)]
    batch = worker.run_batch(jobs)
    singles = [worker.run_job(j) for j in jobs]
    assert [b[1] for b in batch] == [100, 101, 102, 103]
    from PIL import Image
    import numpy as np
    for (pb, _), (ps, _) in zip(batch, singles):
        a = np.asarray(Image.open(io.BytesIO(pb))).astype(int)
        b = np.asarray(Image.open(io.BytesIO(ps))).astype(int)
        assert np.abs(a - b).max() <= 1
    cfg = Mock()
    cfg.config.model_root = os.environ["MODEL_ROOT"]
    mode = Mock(model=os.environ["MODEL"], model_path="x", loras=[])
    cfg.get_mode.return_value = mode
    cfg.get_default_mode.return_value = "tiny"
    reg = Mock()
    reg.get_used_vram.return_value = 0
    pool = WorkerPool(queue_max=16, worker_factory=lambda worker_id: worker, mode_config=cfg,
                      registry=reg, num_workers=1, max_batch=4)
    futs = [pool.submit_job(GenerationJob(req=j.req)) for j in jobs]
    res = [f.result(timeout=120) for f in futs]
    assert [r[1] for r in res] == [100, 101, 102, 103] and all(r[0][:4] == b"\x89PNG" for r in res)
    pool._workers = []        # the fixture owns the worker
    pool.shutdown()


def test_style_lora_levels_and_reset(tmp_path):
    """Style LoRAs through the worker (reference `backends/cuda_worker.py:123-196, 215-232`):
    registry preload, level -> adapter weight, exclusive selection, reset after every job; the
    styled image equals the un-styled pipeline run on weights merged offline in fp32."""
    import json
    import numpy as np
    from PIL import Image
    from safetensors.torch import load_file, save_file
    from dreamlab_b200 import synthetic as S
    from dreamlab_b200.engine import LCMPipelineB200
    from dreamlab_b200.lora import lora_weight_deltas
    from backends.b200_worker import B200Worker, unet_cfg_from_json, vae_cfg_from_json
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from test_lora import MODS, _kohya_lora
    root = tmp_path / "models"
    ucfg = UNetConfig.tiny()
    ucfg.cross_attention_dim = 768
    mdir = S.write_model_dir(str(root / "tiny-lcm"), ucfg, VAEConfig.tiny())
    shapes = S.unet_shapes(ucfg)
    lora = _kohya_lora(shapes, MODS, r=4, alpha=4.0)
    lora = {k: v * 3.0 for k, v in lora.items()}                      # strong enough to see
    save_file({k: v.contiguous() for k, v in lora.items()}, str(tmp_path / "style.safetensors"))
    reg = {"ink": {"title": "Ink", "lora_path": str(tmp_path / "style.safetensors"), "adapter_name": "style_ink",
                   "levels": [0.5, 1.0], "required_cross_attention_dim": 768},
           "xl_only": {"lora_path": str(tmp_path / "style.safetensors"), "levels": [1.0],
                       "required_cross_attention_dim": 2048}}
    (tmp_path / "styles.json").write_text(json.dumps(reg))
    old = {k: os.environ.get(k) for k in ("MODEL_ROOT", "MODEL", "B200_STYLES")}
    os.environ.update(MODEL_ROOT=str(root), MODEL="tiny-lcm", B200_STYLES=str(tmp_path / "styles.json"))
    try:
        w = B200Worker(worker_id=0)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    assert w._style_loaded == {"style_ink": True, "style_xl_only": False}

    def styled(style, level, seed=5):
        j = job(seed=seed)
        j.req.style_lora = SimpleNamespace(style=style, level=level)
        return np.asarray(Image.open(io.BytesIO(w.run_job(j)[0]))).astype(int)

    plain = styled(None, 0)
    lvl1, lvl2, lvl9 = styled("ink", 1), styled("ink", 2), styled("ink", 9)
    assert np.abs(lvl1 - plain).max() > 2 and np.abs(lvl2 - lvl1).max() > 2
    assert np.array_equal(lvl9, lvl2)                                 # level clamps to the ladder
    assert np.array_equal(styled(None, 0), plain)                     # reset: no state bleed
    assert np.array_equal(styled("unknown", 2), plain) and np.array_equal(styled("xl_only", 1), plain)
    # same request on a pipeline built from offline-merged weights (fp32 merge, then packed)
    unet_sd = load_file(os.path.join(mdir, "unet", "diffusion_pytorch_model.safetensors"))
    vae_sd = load_file(os.path.join(mdir, "vae", "diffusion_pytorch_model.safetensors"))
    deltas, _ = lora_weight_deltas(shapes, lora)
    merged = {k: (v.float() + 1.0 * deltas[k] if k in deltas else v) for k, v in unet_sd.items()}
    ref = LCMPipelineB200(merged, unet_cfg_from_json(json.load(open(os.path.join(mdir, "unet", "config.json")))),
                          vae_sd, vae_cfg_from_json(json.load(open(os.path.join(mdir, "vae", "config.json")))), "cuda:0")
    ref.unet.fold_ln = False          # like the worker's UNet: style adapters keep the standalone LayerNorm kernel
    assert w.pipe.unet.fold_ln is False
    lat, noise = w._draw(5, 16, 16, 2)
    pe = w._text.encode(["a cat"])
    img = ref.generate(pe, lat, torch.stack(noise), 2, torch.tensor([1.0])).cpu().numpy()[0].astype(int)
    # the worker adds the delta to the fp32 copy of the PACKED (bf16) weight, the offline merge to the fp16 file weight:
    # one extra rounding -> a few pixels move by up to a few u8 steps
    d = np.abs(img - lvl2)
    assert d.max() <= 4 and (d > 1).mean() < 1e-2, (d.max(), (d > 1).mean())


def test_worker_runs_its_text_tower_on_the_device(tmp_path):
    """A model dir with `text_encoder/` (transformers layout): the worker encodes prompts with the
    on-device CLIP tower (no `transformers` model in the request path) and the embeddings match
    transformers' CLIPTextModel on the same ids."""
    from safetensors.torch import save_file
    from transformers import CLIPTextConfig, CLIPTextModel
    from dreamlab_b200 import synthetic as S
    from dreamlab_b200.clip import CLIPTextB200
    from backends.b200_worker import B200Worker, _hash_tokens
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    root = tmp_path / "models"
    ucfg = UNetConfig.tiny()
    ucfg.cross_attention_dim = 768
    mdir = S.write_model_dir(str(root / "tiny-lcm"), ucfg, VAEConfig.tiny())
    torch.manual_seed(0)
    ccfg = CLIPTextConfig(hidden_size=768, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=12)
    clip = CLIPTextModel(ccfg).eval()
    os.makedirs(os.path.join(mdir, "text_encoder"))
    with open(os.path.join(mdir, "text_encoder", "config.json"), "w") as f:
        f.write(ccfg.to_json_string())
    save_file({k: v.contiguous() for k, v in clip.state_dict().items()},
              os.path.join(mdir, "text_encoder", "model.safetensors"))
    old = {k: os.environ.get(k) for k in ("MODEL_ROOT", "MODEL")}
    os.environ.update(MODEL_ROOT=str(root), MODEL="tiny-lcm")
    try:
        w = B200Worker(worker_id=0)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    assert isinstance(w._text.model, CLIPTextB200) and w.pipe.text_encoder is w._text.model
    prompts = ["a cat", "a much longer prompt about mountains, rivers and a small red house"]
    got = w._text.encode(prompts).cpu()
    with torch.no_grad():
        ref = clip(_hash_tokens(prompts)).last_hidden_state
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err <= 2e-2, err
    # the tower runs as one CUDA-graph replay per batch bucket (static ids in, static states out): a second call
    # through the same graph, a padded bucket (3 prompts in the 4-bucket) and the eager path all agree
    assert len(w._text.model._graphs) == 1
    three = ["x", prompts[1], "third prompt"]
    g3 = w._text.encode(three).cpu()
    assert torch.equal(w._text.encode(prompts).cpu(), got) and len(w._text.model._graphs) == 2
    eager = w._text.model.forward(_hash_tokens(three))["last_hidden_state"].float().cpu()
    assert torch.equal(g3, eager)
    assert torch.equal(g3[1], got[1])                                  # a prompt's embedding does not depend on its batch
    a, b = w.run_job(job(prompt=prompts[0], seed=3)), w.run_job(job(prompt=prompts[1], seed=3))
    assert a[0][:4] == b"\x89PNG" and a[0] != b[0]                   # the prompt conditions the image
    assert w.run_job(job(prompt=prompts[0], seed=3))[0] == a[0]


def test_yume_style_quick_job(worker):
    """The dream loop's candidates (`yume/dream_worker.py:262-299`): ad-hoc request objects without
    `style_lora`, 64x64, ONE step; `run_job_array` returns the same pixels without the PNG round trip."""
    import numpy as np
    from PIL import Image
    req = SimpleNamespace(prompt="dream of electric sheep", size="64x64", num_inference_steps=1,
                          guidance_scale=1.0, seed=11)                 # no style_lora attribute
    j = SimpleNamespace(req=req, fut=None, submitted_at=0.0)
    png, seed = worker.run_job(j)
    arr, seed2 = worker.run_job_array(j)
    assert seed == seed2 == 11 and arr.shape == (64, 64, 3) and arr.dtype == np.uint8
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(png))), arr)


def test_gpu_png_mode_same_pixels_and_deterministic(worker, monkeypatch):
    """B200_PNG=gpu: the PNG file is assembled on the device (csrc/png.cu).  The contract the reference's tests
    pin (`tests/test_sdxl_worker.py:139-198`): PNG magic, seed echo, same seed => byte-identical file; plus it
    decodes to exactly the pixels of the default (PIL) path and equals the oracle's stored-deflate file."""
    import numpy as np
    from PIL import Image
    from oracle.png import png_stored
    ref_png, _ = worker.run_job(job(seed=42))
    arr, _ = worker.run_job_array(job(seed=42))
    monkeypatch.setenv("B200_PNG", "gpu")
    png, seed = worker.run_job(job(seed=42))
    assert png[:8] == b"\x89PNG\r\n\x1a\n" and seed == 42
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(png))), np.asarray(Image.open(io.BytesIO(ref_png))))
    assert png == png_stored(arr)
    assert worker.run_job(job(seed=42))[0] == png
    batch = worker.run_batch([job(prompt=f"p{i}", seed=200 + i) for i in range(3)], with_latents=True)
    assert all(b[0][:8] == b"\x89PNG\r\n\x1a\n" and len(b[2]) == 512 for b in batch)
    monkeypatch.setenv("B200_PNG", "jpeg")
    with pytest.raises(RuntimeError, match="B200_PNG"):
        worker.run_job(job(seed=42))


def test_two_real_workers_capture_and_serve_concurrently(tmp_path):
    """WorkerPool with TWO real B200Workers in one process (on a one-GPU box both land on cuda:0): every
    worker captures its CUDA graphs lazily on first use of a geometry, so mixed batch sizes arriving at
    once make the two threads warm up / capture / replay concurrently (capture in "global" error
    mode made the other thread's CUDA calls fail).  Every request must come back, and equal the same request
    served alone."""
    from dreamlab_b200 import synthetic as S
    from oracle.unet import UNetConfig
    from oracle.vae import VAEConfig
    from backends.worker_factory import create_cuda_worker
    from backends.worker_pool import GenerationJob, WorkerPool
    from unittest.mock import Mock
    ucfg = UNetConfig.tiny()
    ucfg.cross_attention_dim = 768
    S.write_model_dir(str(tmp_path / "tiny-lcm"), ucfg, VAEConfig.tiny())
    old = {k: os.environ.get(k) for k in ("MODEL_ROOT", "MODEL", "CUDA_DEVICE")}
    os.environ["MODEL_ROOT"], os.environ["MODEL"] = str(tmp_path), "tiny-lcm"
    os.environ.pop("CUDA_DEVICE", None)
    try:
        cfg = Mock()
        cfg.config.model_root = str(tmp_path)
        cfg.get_mode.return_value = Mock(model="tiny-lcm", model_path="x", loras=[])
        cfg.get_default_mode.return_value = "tiny"
        reg = Mock()
        reg.get_used_vram.return_value = 0
        pool = WorkerPool(queue_max=256, worker_factory=create_cuda_worker, mode_config=cfg, registry=reg,
                          num_workers=2, max_batch=4)
        assert len(pool._workers) == 2 and pool._workers[0] is not pool._workers[1]
        reqs = []
        for rnd in range(3):                                   # bursts of different geometry / batch size
            for i, (size, steps) in enumerate([("64x64", 2)] * 3 + [("128x128", 2)] * 5 + [("64x128", 1)] * 2):
                reqs.append(job(prompt=f"r{rnd}-{i}", size=size, steps=steps, seed=1000 * rnd + i).req)
        futs = [pool.submit_job(GenerationJob(req=r)) for r in reqs]
        res = [f.result(timeout=300) for f in futs]
        assert all(png[:8] == b"\x89PNG\r\n\x1a\n" for png, _ in res)
        assert [s for _, s in res] == [r.seed for r in reqs]
        solo = pool._workers[0]
        import numpy as np
        from PIL import Image
        for k in (0, 4, 9, 17, 29):
            png, _ = solo.run_job(SimpleNamespace(req=reqs[k]))
            a = np.asarray(Image.open(io.BytesIO(png))).astype(int)
            b = np.asarray(Image.open(io.BytesIO(res[k][0]))).astype(int)
            assert np.abs(a - b).max() <= 1, k                 # batch-invariance bar of test_batch_invariance
        pool.shutdown()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_latent_only_jobs_match_the_image_pass(worker):
    """`run_job_latents` / `run_batch(latents_only=True)`: the denoise loop without VAE decode or PNG (Yume-style
    candidate scoring).  The latents are those the image pass decodes: pooled to 8x8 they equal the 512 bytes
    `run_job_with_latents` returns for the same request."""
    import numpy as np
    j = job(seed=21, size="128x128", steps=2)
    lat, seed = worker.run_job_latents(j)
    assert seed == 21 and lat.shape == (4, 16, 16) and lat.dtype == np.float32
    _, _, lat_bytes = worker.run_job_with_latents(j)
    pooled = torch.nn.functional.adaptive_avg_pool2d(torch.from_numpy(lat)[None], (8, 8)).to(torch.float16)
    ref = np.frombuffer(lat_bytes, dtype=np.float16).reshape(1, 4, 8, 8)
    assert np.abs(pooled.numpy().astype(np.float32) - ref.astype(np.float32)).max() <= 2e-3
    many = worker.run_batch([job(prompt=f"c{i}", seed=300 + i, size="64x64", steps=1) for i in range(5)],
                            latents_only=True)
    assert [s for _, s in many] == [300, 301, 302, 303, 304] and all(a.shape == (4, 8, 8) for a, _ in many)
    solo, _ = worker.run_job_latents(job(prompt="c2", seed=302, size="64x64", steps=1))
    assert np.abs(solo - many[2][0]).max() <= 2e-2 * np.abs(solo).max()


@pytest.mark.parametrize("png", ["pil", "gpu"])
def test_deferred_batches_overlap_and_match_the_synchronous_results(worker, png, monkeypatch):
    """run_batch(deferred=True) returns while the batch is still on the GPU (the result copy is only enqueued; each
    thunk waits on the batch's event): four batches are enqueued back to back — at most B200_BATCHES_IN_FLIGHT
    outstanding — and resolved afterwards in reverse order; every result equals the synchronous run_batch of the same
    jobs (static graph buffers, pinned staging and PNG buffers of neighbouring batches do not alias)."""
    monkeypatch.setenv("B200_PNG", png)
    batches = [[job(prompt=f"p{b}-{i}", seed=100 * b + i) for i in range(3 + b % 2)] for b in range(4)]
    ref = [worker.run_batch(bt) for bt in batches]
    pending = [worker.run_batch(bt, deferred=True) for bt in batches]
    assert len(worker._in_flight) <= 2
    assert all(callable(getattr(t, "wait", None)) for th in pending for t in th)
    got = [None] * len(batches)
    for b in reversed(range(len(batches))):
        got[b] = [t() for t in pending[b]]
    assert got == ref
    worker.drain()
    assert len(worker._in_flight) == 0
