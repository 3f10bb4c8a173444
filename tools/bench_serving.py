"""End-to-end serving throughput of ONE b200 worker behind the WorkerPool (requests in, PNG bytes
out): N queued 512x512 4-step requests, micro-batched, with the PNG encoding deferred to encoder
threads vs done on the worker thread.  python tools/bench_serving.py [n_requests]"""
import os
import sys
import tempfile
import time
from types import SimpleNamespace
from unittest.mock import Mock

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamlab_b200 import synthetic as S

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
root = tempfile.mkdtemp()
S.write_model_dir(os.path.join(root, "sd15-lcm"))
os.environ.update(MODEL_ROOT=root, MODEL="sd15-lcm")
from backends.b200_worker import B200Worker
from backends.worker_pool import GenerationJob, WorkerPool, reset_worker_pool

cfg = Mock()
cfg.config.model_root = root
cfg.get_mode.return_value = Mock(model="sd15-lcm", model_path="x", loras=[])
cfg.get_default_mode.return_value = "sd15"
reg = Mock()
reg.get_used_vram.return_value = 0
for deferred in (True, False):
    B200Worker.supports_deferred = deferred
    reset_worker_pool()
    pool = WorkerPool(queue_max=2 * N, worker_factory=lambda worker_id: B200Worker(worker_id=worker_id),
                      mode_config=cfg, registry=reg, num_workers=1, max_batch=16)

    def submit(n, seed0):
        return [pool.submit_job(GenerationJob(req=SimpleNamespace(
            prompt=f"p{i}", size="512x512", num_inference_steps=4, guidance_scale=1.0, seed=seed0 + i)))
            for i in range(n)]
    [f.result(timeout=300) for f in submit(32, 0)]          # warm-up: graph capture, encoder threads
    t0 = time.perf_counter()
    outs = [f.result(timeout=600) for f in submit(N, 1000)]
    dt = time.perf_counter() - t0
    kb = sum(len(o[0]) for o in outs) / len(outs) / 1024
    print(f"deferred_png={deferred}: {N} requests in {dt:.2f} s = {N / dt:.1f} img/s end to end "
          f"(PNG {kb:.0f} KB avg, {os.cpu_count()} host cores)", flush=True)
    pool.shutdown()
