"""The reference's worker class names, bound to the B200 workers.

Reference `backends/cuda_worker.py` defines `DiffusersCudaWorker` (`:26-304`) and `DiffusersSDXLCudaWorker`
(`:307-614`); its factory, its server code and its own tests (`tests/test_worker_factory.py:133-159`) refer to the
workers by these names.  Here they ARE `B200Worker` / `B200SDXLWorker` — same constructor (`worker_id`), environment,
attributes, results and errors (see `backends/b200_worker.py`); nothing of diffusers is behind them.
"""
from backends.b200_worker import B200SDXLWorker, B200Worker

DiffusersCudaWorker = B200Worker
DiffusersSDXLCudaWorker = B200SDXLWorker

# what the two names were bound to at import: the factory uses it to see whether a caller rebound one of them
_ORIGINALS = {"DiffusersCudaWorker": B200Worker, "DiffusersSDXLCudaWorker": B200SDXLWorker}

__all__ = ["DiffusersCudaWorker", "DiffusersSDXLCudaWorker"]
