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


def test_sdxl_is_reported_unsupported_outside_reference_tree():
    with patch("backends.worker_factory.detect_worker_type", return_value="sdxl"), \
            patch.dict(sys.modules, {"backends.cuda_worker": None}):
        with pytest.raises(RuntimeError, match="SDXL"):
            create_cuda_worker(worker_id=0)


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
