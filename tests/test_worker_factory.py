"""Factory: dim -> worker-type table, env handling, b200 dispatch (strategy of the reference's
`tests/test_worker_factory.py`: detection and worker classes are patched, nothing loads)."""
import json
import os
import sys
from types import SimpleNamespace
from unittest.mock import Mock, patch

import pytest

from backends.worker_factory import create_cuda_worker, detect_worker_type


def _info(dim):
    return SimpleNamespace(cross_attention_dim=dim, confidence=0.9, variant=SimpleNamespace(value="x"))


@pytest.mark.parametrize("dim,want", [(768, "sd15"), (1024, "sd15"), (2048, "sdxl"), (1280, "sdxl")])
def test_dim_table(dim, want):
    with patch.dict(os.environ, {"MODEL_ROOT": "/models", "MODEL": " m.safetensors "}), \
            patch("os.path.exists", return_value=True), \
            patch("backends.worker_factory._detect_model", return_value=_info(dim)):
        assert detect_worker_type() == want


@pytest.mark.parametrize("dim", [512, 0, None, 4096])
def test_unsupported_dim_raises(dim):
    with patch.dict(os.environ, {"MODEL_ROOT": "/models", "MODEL": "m"}), \
            patch("os.path.exists", return_value=True), \
            patch("backends.worker_factory._detect_model", return_value=_info(dim)):
        with pytest.raises(RuntimeError, match="Unsupported cross_attention_dim"):
            detect_worker_type()


def test_env_errors():
    with patch.dict(os.environ, {}, clear=True):
        with pytest.raises(RuntimeError, match="MODEL_ROOT"):
            detect_worker_type()
    with patch.dict(os.environ, {"MODEL_ROOT": "/models"}, clear=True):
        with pytest.raises(RuntimeError, match="MODEL environment"):
            detect_worker_type()
    with patch.dict(os.environ, {"MODEL_ROOT": "/models", "MODEL": "gone"}), \
            patch("os.path.exists", return_value=False):
        with pytest.raises(RuntimeError, match="Model not found"):
            detect_worker_type()


def test_builtin_detector_reads_diffusers_layout(tmp_path):
    d = tmp_path / "lcm"
    (d / "unet").mkdir(parents=True)
    (d / "unet" / "config.json").write_text(json.dumps({"cross_attention_dim": 768}))
    with patch.dict(os.environ, {"MODEL_ROOT": str(tmp_path), "MODEL": "lcm"}), \
            patch.dict(sys.modules, {"utils.model_detector": None}):
        assert detect_worker_type() == "sd15"


def test_sd15_gets_b200_worker_with_worker_id_kwarg():
    fake = Mock()
    with patch("backends.worker_factory.detect_worker_type", return_value="sd15"), \
            patch("backends.b200_worker.B200Worker", fake):
        w = create_cuda_worker(worker_id=3)
    fake.assert_called_once_with(worker_id=3)
    assert w is fake.return_value


def test_sdxl_gets_b200_sdxl_worker_with_worker_id_kwarg():
    """Reference `backends/worker_factory.py:91-94` (sdxl -> SDXL worker class, keyword ctor)."""
    fake = Mock()
    with patch("backends.worker_factory.detect_worker_type", return_value="sdxl"), \
            patch("backends.b200_worker.B200SDXLWorker", fake):
        w = create_cuda_worker(worker_id=5)
    fake.assert_called_once_with(worker_id=5)
    assert w is fake.return_value


def test_sdxl_worker_contract_errors_without_loading():
    """Same env errors as `DiffusersSDXLCudaWorker.__init__` (reference `cuda_worker.py:343-346`)."""
    from backends.b200_worker import B200SDXLWorker
    with patch.dict(os.environ, {}, clear=True):
        with pytest.raises(RuntimeError, match="MODEL_ROOT is required for SDXL CUDA worker"):
            B200SDXLWorker(worker_id=0)
    with patch.dict(os.environ, {"MODEL_ROOT": "/models"}, clear=True):
        with pytest.raises(RuntimeError, match="MODEL is required for SDXL CUDA worker"):
            B200SDXLWorker(worker_id=0)


def test_unet_cfg_from_json_sdxl():
    from backends.b200_worker import unet_cfg_from_json
    c = unet_cfg_from_json({"block_out_channels": [320, 640, 1280],
                            "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
                            "attention_head_dim": [5, 10, 20], "transformer_layers_per_block": [1, 2, 10],
                            "cross_attention_dim": 2048, "use_linear_projection": True,
                            "addition_embed_type": "text_time", "addition_time_embed_dim": 256,
                            "projection_class_embeddings_input_dim": 2816})
    assert c.down_attn == (False, True, True) and c.attention_head_dim == (5, 10, 20)
    assert c.transformer_layers_per_block == (1, 2, 10) and c.use_linear_projection
    assert c.time_cond_proj_dim is None and c.addition_embed_type == "text_time"
    d = unet_cfg_from_json({})
    assert d.attention_head_dim == 8 and d.transformer_layers_per_block == () and not d.use_linear_projection


def test_b200_worker_contract_errors_without_loading():
    from backends.b200_worker import B200Worker, parse_size
    assert parse_size("512x768") == (512, 768) and parse_size("512X512") == (512, 512)
    for bad in ("512", "axb", "512x", None, "1x2x3"):
        with pytest.raises(RuntimeError, match="Invalid size"):
            parse_size(bad)
    with patch.dict(os.environ, {}, clear=True):
        with pytest.raises(RuntimeError, match="MODEL_ROOT is required"):
            B200Worker(worker_id=0)
    with patch.dict(os.environ, {"MODEL_ROOT": "/models"}, clear=True):
        with pytest.raises(RuntimeError, match="MODEL is required"):
            B200Worker(worker_id=0)


def test_png_compress_level_knob(monkeypatch):
    """B200_PNG_COMPRESS_LEVEL: unset = the reference's exact PIL call; any level decodes to the
    same pixels; a bad value fails loudly."""
    import io
    import numpy as np
    from PIL import Image
    from backends.b200_worker import _encode_png
    arr = np.random.default_rng(0).integers(0, 256, (64, 48, 3), dtype=np.uint8)
    monkeypatch.delenv("B200_PNG_COMPRESS_LEVEL", raising=False)
    ref = io.BytesIO()
    Image.fromarray(arr).save(ref, format="PNG")
    assert _encode_png(arr) == ref.getvalue()
    for lvl in ("0", "1", "9"):
        monkeypatch.setenv("B200_PNG_COMPRESS_LEVEL", lvl)
        png = _encode_png(arr)
        assert png[:8] == b"\x89PNG\r\n\x1a\n"
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(png))), arr)
    monkeypatch.setenv("B200_PNG_COMPRESS_LEVEL", "12")
    with pytest.raises(RuntimeError, match="0..9"):
        _encode_png(arr)
