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
