"""Worker factory with model-type detection — B200 edition.

Keeps the reference's entry points (`backends/worker_factory.py:17-100`):
`detect_worker_type() -> "sd15" | "sdxl"` from the `cross_attention_dim` of the model named by
`MODEL_ROOT`/`MODEL` (768/1024 -> sd15, 1280/2048 -> sdxl, anything else raises) and
`create_cuda_worker(worker_id)`.  The only behavioural change: SD1.5-class models now get the
B200-native `B200Worker` instead of `DiffusersCudaWorker`, SDXL-class models `B200SDXLWorker`
instead of `DiffusersSDXLCudaWorker` (`backends/cuda_worker.py` keeps the reference's two names, bound to them).  There is no switch back to a diffusers worker and no other
backend: this package is the B200 path only.
"""
from __future__ import annotations

import json
import logging
import os
from types import SimpleNamespace
from typing import TYPE_CHECKING

if TYPE_CHECKING:
    from backends.base import PipelineWorker

logger = logging.getLogger(__name__)

_SDXL_DIMS = (2048, 1280)      # SDXL base / refiner
_SD15_DIMS = (768, 1024)       # SD1.x / SD2.x


def _builtin_detect(model_path: str):
    """Stand-in for the reference's `utils.model_detector.detect_model`: diffusers-layout dirs
    (unet/config.json) and single-file .safetensors checkpoints (header only).  For single files it is used even
    inside the reference tree: the reference's SafetensorsDetector reads `attn2.to_k.weight.shape[0]`, the
    OUTPUT dim of that Linear (320 / 640 / 1280), not cross_attention_dim."""
    if os.path.isfile(model_path):
        from dreamlab_b200.single_file import sniff
        info = sniff(model_path)
        return SimpleNamespace(cross_attention_dim=info["cross_attention_dim"], confidence=1.0,
                               variant=SimpleNamespace(value=info["variant"]))
    cfg = os.path.join(model_path, "unet", "config.json")
    if not os.path.exists(cfg):
        raise RuntimeError(f"cannot inspect {model_path}: no unet/config.json "
                           "(single-file checkpoints need the reference's utils.model_detector)")
    with open(cfg) as f:
        dim = json.load(f).get("cross_attention_dim")
    return SimpleNamespace(cross_attention_dim=dim, confidence=1.0,
                           variant=SimpleNamespace(value="sdxl" if dim in _SDXL_DIMS else "sd15"))


def _detect_model(model_path: str):
    if os.path.isfile(model_path) and model_path.lower().endswith(".safetensors"):
        return _builtin_detect(model_path)
    try:
        from utils.model_detector import detect_model      # reference tree
    except ImportError:
        return _builtin_detect(model_path)
    return detect_model(model_path)


def detect_worker_type() -> str:
    model_root = os.environ.get("MODEL_ROOT", "").strip()
    model_name = os.environ.get("MODEL", "").strip()
    if not model_root:
        raise RuntimeError("MODEL_ROOT environment variable is required")
    if not model_name:
        raise RuntimeError("MODEL environment variable is required")
    model_path = os.path.join(model_root, model_name)
    if not os.path.exists(model_path):
        raise RuntimeError(f"Model not found at: {model_path}")
    try:
        info = _detect_model(model_path)
        dim = info.cross_attention_dim
        logger.info("[ModelDetection] %s: cross_attention_dim=%s", model_path, dim)
        if dim in _SDXL_DIMS:
            return "sdxl"
        if dim in _SD15_DIMS:
            return "sd15"
        raise RuntimeError(
            f"Unsupported cross_attention_dim: {dim}. Expected 768 (SD1.5), 1024 (SD2.x), "
            f"1280 (SDXL Refiner), or 2048 (SDXL Base)")
    except Exception as e:
        logger.error("[ModelDetection] Failed to detect model: %s", e)
        raise RuntimeError(f"Model detection failed: {e}")


def _worker_class(worker_type: str):
    """The class behind a worker type, looked up at call time under both of its names: the reference's
    (`backends.cuda_worker.DiffusersCudaWorker` / `DiffusersSDXLCudaWorker`) and this package's
    (`backends.b200_worker.B200Worker` / `B200SDXLWorker`).  A caller that rebound the reference name (its own tests
    do, `tests/test_worker_factory.py:133-159`) gets what it installed."""
    import backends.b200_worker as bw
    ref_name, name = (("DiffusersSDXLCudaWorker", "B200SDXLWorker") if worker_type == "sdxl"
                      else ("DiffusersCudaWorker", "B200Worker"))
    try:
        import backends.cuda_worker as cw
    except ImportError:                       # the reference's own cuda_worker.py without diffusers installed
        return getattr(bw, name)
    originals = getattr(cw, "_ORIGINALS", None)
    if originals is None:                     # the reference's own cuda_worker.py: its diffusers workers are not ours
        return getattr(bw, name)
    cls = getattr(cw, ref_name)
    return cls if cls is not originals[ref_name] else getattr(bw, name)


def create_cuda_worker(worker_id: int) -> "PipelineWorker":
    worker_type = detect_worker_type()
    cls = _worker_class(worker_type)
    worker = cls(worker_id=worker_id)
    logger.info("[WorkerFactory] Created %s (worker %d)", getattr(cls, "__name__", cls), worker_id)
    return worker
