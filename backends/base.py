"""Worker interface — the drop-in boundary (mirrors reference `backends/base.py:8-58`).

`PipelineWorker` is what `WorkerPool` drives (`job.execute(worker)` -> `worker.run_job(job)`,
reference `backends/worker_pool.py:84-88`).  A request object is duck-typed: workers read
`req.prompt`, `req.size` ("WIDTHxHEIGHT"), `req.num_inference_steps`, `req.guidance_scale`,
`req.seed` (optional) and `getattr(req, "style_lora", None)`.
"""
from __future__ import annotations

import os
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Any, Optional, Protocol, Tuple, runtime_checkable


@dataclass(frozen=True)
class StyleLora:
    style: Optional[str] = None      # e.g. "papercut"
    level: int = 0                   # 0 = off, 1..N preset index


@dataclass
class GenSpec:
    prompt: str
    size: str
    steps: int
    cfg: float
    seed: Optional[int] = None
    style_lora: StyleLora = field(default_factory=StyleLora)


@dataclass
class Job:
    req: Any                         # GenerateRequest-like
    fut: Future
    submitted_at: float


@runtime_checkable
class PipelineWorker(Protocol):
    worker_id: int

    def run_job(self, spec) -> Tuple[bytes, int]:
        """Return (png_bytes, seed_used)."""

    def run_job_with_latents(self, spec) -> Tuple[bytes, int, bytes]:
        """Return (png_bytes, seed_used, latents_bytes): latents_bytes is the final latent
        average-pooled to [1,4,8,8] (NCHW), little-endian float16 — 512 bytes."""


@dataclass(frozen=True)
class ModelPaths:
    root: str

    @property
    def scheduler_config(self) -> str:
        return os.path.join(self.root, "scheduler", "scheduler_config.json")

    @property
    def text_encoder(self) -> str:
        return os.path.join(self.root, "text_encoder")

    @property
    def unet(self) -> str:
        return os.path.join(self.root, "unet")

    @property
    def vae_decoder(self) -> str:
        return os.path.join(self.root, "vae_decoder")
