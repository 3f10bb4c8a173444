"""The C-ABI library builds, loads and exports exactly what include/dreamlab_b200.h declares
(no compute here: there is no GPU in the CPU test container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dreamlab_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import dreamlab_b200.lib as L
    return L


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 15
    l = lib.load()
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == names
    assert l.dl_abi_version() == 3


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.load().dl_device_sm_count() == -1
    assert b"no CUDA device" in lib.load().dl_last_error()
    with pytest.raises(RuntimeError):
        lib.require_cuda()
    from dreamlab_b200.engine import UNetB200
    with pytest.raises(RuntimeError):
        UNetB200({}, None, "cuda:0")


def test_binary_is_sm100a_tcgen05():
    """SASS evidence that the hot kernels are Blackwell-native (UTCHMMA = tcgen05.mma,
    UTMALDG = TMA, LDTM/STTM = tcgen05.ld/st)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "stable-diffusion-1.5-lcm-onnx-rknn2_b200", "libdreamlab_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnem in sass, mnem


def test_product_never_imports_oracle():
    pk = os.path.join(ROOT, "stable-diffusion-1.5-lcm-onnx-rknn2_b200")
    for base in (pk, os.path.join(ROOT, "backends")):
        for dp, _, fs in os.walk(base):
            for f in fs:
                if f.endswith(".py"):
                    txt = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_wait_for_driver_returns_at_once_without_a_gpu_node(monkeypatch):
    """lib.wait_for_driver() only retries cuInit when a GPU device node exists; in the CPU container it must not
    delay the loud 'no CUDA device' failure of the product path."""
    import glob
    import time
    from dreamlab_b200 import lib
    monkeypatch.setattr(glob, "glob", lambda pat: [])
    t0 = time.monotonic()
    lib.wait_for_driver(timeout_s=30.0)
    assert time.monotonic() - t0 < 1.0
