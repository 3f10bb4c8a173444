"""The reference's OWN tests for the boundary of the hot path, run against this repo's `backends/`.

`/root/reference/tests/test_worker_pool.py` and `test_worker_factory.py` exercise `backends.worker_pool`
(`WorkerPool`, the job types, FIFO, mode switches, shutdown) and `backends.worker_factory`
(`detect_worker_type`, `create_cuda_worker`, the `backends.cuda_worker` class names) through mocks — exactly the
interface SURVEY.md §8 rows a1 / a15 / a16 / b name.  They are copied to a scratch directory and run unmodified
in a child pytest with this repo first on the path (so `backends` is ours) and the reference tree behind it (its
`utils.model_detector` is what the factory tests patch).  This container only: the GPU box has no reference tree.
"""
import os
import shutil
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
FILES = ("test_worker_pool.py", "test_worker_factory.py")


@pytest.mark.skipif(not all(os.path.exists(os.path.join(REFERENCE, "tests", f)) for f in FILES),
                    reason="reference tree not present (GPU box)")
def test_reference_pool_and_factory_tests_pass_against_our_backends(tmp_path):
    dst = tmp_path / "tests"
    dst.mkdir()
    for f in FILES + ("conftest.py", "__init__.py"):
        src = os.path.join(REFERENCE, "tests", f)
        if os.path.exists(src):
            shutil.copy(src, dst / f)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([REPO, REFERENCE]))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + [f"tests/{f}" for f in FILES],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    # the child really imported OUR backends
    probe = subprocess.run([sys.executable, "-c", "import backends.worker_pool as m; print(m.__file__)"],
                           cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=120)
    assert probe.stdout.strip().startswith(REPO), probe.stdout + probe.stderr
