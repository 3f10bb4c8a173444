"""Where a B200Worker batch spends host time: cProfile over N `run_batch(jobs, deferred=True)` calls + their thunks
(bench.py's model dir and request shape), plus wall time per batch against the graph-replay time.
  python tools/profile_worker_batch.py [batches]"""
import cProfile
import io
import os
import pstats
import sys
import time
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from backends.worker_pool import GenerationJob

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 10
c = dict(bench.CONFIGS["c2"])
os.environ.setdefault("B200_PNG", "gpu")
root, name = bench.bench_model_dir(c, 0, lambda: None)
os.environ["MODEL_ROOT"], os.environ["MODEL"] = root, name
from backends.b200_worker import B200Worker
w = B200Worker(worker_id=0)
B = c["batch"]


def jobs(k):
    return [GenerationJob(req=SimpleNamespace(prompt=f"bench prompt {k * B + i}", size="512x512",
                                              num_inference_steps=c["lcm_steps"], guidance_scale=c["gs"],
                                              seed=k * B + i)) for i in range(B)]


for k in range(3):
    [t() for t in w.run_batch(jobs(k), deferred=True)]
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for k in range(n_batches):
    thunks = w.run_batch(jobs(100 + k), deferred=True)
    out = [t() for t in thunks]
pr.disable()
dt = (time.perf_counter() - t0) / n_batches
print(f"wall per batch of {B}: {dt * 1e3:.2f} ms  ({B / dt:.1f} img/s, one thread, thunks inline)")
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
