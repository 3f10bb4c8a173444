"""Worker pool: the reference's contract (its own `tests/test_worker_pool.py` also passes against
this module) plus what is new here — N workers, micro-batching, barrier switch."""
import queue
import threading
import time
from concurrent.futures import Future
from types import SimpleNamespace
from unittest.mock import Mock

import pytest

from backends.worker_pool import (CustomJob, GenerationJob, JobType, ModeSwitchJob, WorkerPool,
                                  get_worker_pool, reset_worker_pool)


def make_config():
    cfg = Mock()
    cfg.config = Mock()
    cfg.config.model_root = "/models"
    modes = {}
    for name, model in (("sd15-fast", "sd15"), ("sd15-alt", "alt")):
        m = Mock()
        m.name, m.model, m.model_path, m.loras = name, model, f"/models/{model}", []
        modes[name] = m
    cfg.get_mode.side_effect = lambda n: modes[n]
    cfg.get_default_mode.return_value = "sd15-fast"
    return cfg


def make_registry():
    r = Mock()
    r.get_used_vram.return_value = 0
    return r


class FakeWorker:
    """Worker with a real run_batch (what B200Worker offers)."""

    def __init__(self, worker_id, delay=0.0):
        self.worker_id = worker_id
        self.delay = delay
        self.batches = []
        self.lock = threading.Lock()

    def run_batch(self, jobs, with_latents=False):
        time.sleep(self.delay)
        with self.lock:
            self.batches.append(len(jobs))
        return [(f"png-{j.req.prompt}".encode(), self.worker_id) for j in jobs]

    def run_job(self, job):
        return self.run_batch([job])[0]


def req(prompt, size="512x512", steps=4):
    return SimpleNamespace(prompt=prompt, size=size, num_inference_steps=steps, guidance_scale=1.0, seed=1)


@pytest.fixture
def pool_factory():
    pools = []

    def make(**kw):
        reset_worker_pool()
        p = WorkerPool(queue_max=kw.pop("queue_max", 64), mode_config=make_config(),
                       registry=make_registry(), **kw)
        pools.append(p)
        return p
    yield make
    for p in pools:
        p.shutdown()


def test_default_is_one_worker_id0(pool_factory):
    factory = Mock(return_value=Mock(run_job=Mock(return_value="ok")))
    pool = pool_factory(worker_factory=factory, num_workers=None)
    factory.assert_called_once_with(worker_id=0)
    assert pool._current_mode == "sd15-fast" and pool._worker is not None
    assert pool._worker_thread.is_alive() and pool.queue_max == 64
    job = GenerationJob(req=req("a"))
    assert job.job_type == JobType.GENERATION
    assert pool.submit_job(job).result(timeout=5) == "ok"


def test_n_workers_pinned_and_all_used(pool_factory):
    workers = {}

    def factory(worker_id):
        workers[worker_id] = FakeWorker(worker_id, delay=0.05)
        return workers[worker_id]
    pool = pool_factory(worker_factory=factory, num_workers=4, max_batch=1)
    assert sorted(workers) == [0, 1, 2, 3] and len(pool._worker_threads) == 4
    futs = [pool.submit_job(GenerationJob(req=req(str(i)))) for i in range(32)]
    used = {f.result(timeout=10)[1] for f in futs}
    assert used == {0, 1, 2, 3}                      # independent images shard across workers
    assert [f.result()[0] for f in futs] == [f"png-{i}".encode() for i in range(32)]


def test_micro_batching_groups_same_geometry_fifo_prefix(pool_factory):
    w = FakeWorker(0, delay=0.2)
    pool = pool_factory(worker_factory=lambda worker_id: w, num_workers=1, max_batch=8)
    first = pool.submit_job(GenerationJob(req=req("warm")))      # occupies the worker
    time.sleep(0.05)
    futs = [pool.submit_job(GenerationJob(req=req(f"a{i}"))) for i in range(5)]
    futs += [pool.submit_job(GenerationJob(req=req("b", size="768x768")))]
    futs += [pool.submit_job(GenerationJob(req=req(f"c{i}"))) for i in range(2)]
    for f in [first] + futs:
        f.result(timeout=10)
    assert w.batches == [1, 5, 1, 2]                  # never reorders across a different geometry
    assert futs[5].result()[0] == b"png-b"


def test_mode_switch_is_a_barrier_and_recreates_all_workers(pool_factory):
    created = []

    def factory(worker_id):
        created.append(worker_id)
        return FakeWorker(worker_id, delay=0.05)
    pool = pool_factory(worker_factory=factory, num_workers=2, max_batch=1)
    before = [pool.submit_job(GenerationJob(req=req(f"x{i}"))) for i in range(4)]
    sw = pool.switch_mode("sd15-alt")
    after = [pool.submit_job(GenerationJob(req=req(f"y{i}"))) for i in range(4)]
    assert sw.result(timeout=10) == {"mode": "sd15-alt", "status": "switched"}
    for f in before + after:
        f.result(timeout=10)
    assert created == [0, 1, 0, 1] and pool.get_current_mode() == "sd15-alt"
    assert pool.switch_mode("sd15-alt").result(timeout=5)["status"] == "already_loaded"
    with pytest.raises(KeyError):
        pool.switch_mode("nope")


def test_errors_propagate_and_pool_keeps_serving(pool_factory):
    w = Mock()
    w.run_job.side_effect = [RuntimeError("Invalid size 'x', expected 'WIDTHxHEIGHT'"), "fine"]
    pool = pool_factory(worker_factory=lambda worker_id: w, num_workers=1)
    f1 = pool.submit_job(GenerationJob(req=req("bad", size="x")))
    with pytest.raises(RuntimeError, match="Invalid size"):
        f1.result(timeout=5)
    assert pool.submit_job(GenerationJob(req=req("ok"))).result(timeout=5) == "fine"
    assert pool.submit_job(CustomJob(handler=lambda a, b=0: a + b, args=(1,), kwargs={"b": 2})).result(timeout=5) == 3


def test_queue_full_and_shutdown(pool_factory):
    gate = threading.Event()
    w = Mock()
    w.run_job.side_effect = lambda job: (gate.wait(5), "done")[1]
    pool = pool_factory(worker_factory=lambda worker_id: w, num_workers=1, queue_max=2)
    futs = [pool.submit_job(GenerationJob(req=req("0")))]
    time.sleep(0.4)                                   # the worker thread took job 0
    futs += [pool.submit_job(GenerationJob(req=req(str(i)))) for i in (1, 2)]
    with pytest.raises(queue.Full):
        pool.submit_job(GenerationJob(req=req("3")))
    gate.set()
    assert [f.result(timeout=5) for f in futs] == ["done"] * 3
    pool.shutdown()
    assert pool._worker is None and pool.get_queue_size() == 0


def test_singleton_first_call_wins():
    reset_worker_pool()
    f = Mock(return_value=Mock())
    p1 = get_worker_pool(worker_factory=f, mode_config=make_config(), registry=make_registry())
    p2 = get_worker_pool(worker_factory=Mock(), mode_config=make_config(), registry=make_registry())
    assert p1 is p2
    reset_worker_pool()


class DeferredWorker(FakeWorker):
    """Worker offering deferred results (B200Worker.run_batch(..., deferred=True))."""
    supports_deferred = True

    def __init__(self, worker_id, encode_s=0.15):
        super().__init__(worker_id)
        self.encode_s = encode_s
        self.gpu_done = []

    def run_batch(self, jobs, with_latents=False, deferred=False):
        with self.lock:
            self.batches.append(len(jobs))
            self.gpu_done.append(time.perf_counter())

        def finish(j):
            if j.req.prompt == "boom":
                raise ValueError("encode failed")
            time.sleep(self.encode_s)                      # PNG compression stand-in
            return (f"png-{j.req.prompt}".encode(), self.worker_id)
        if deferred:
            import functools
            return [functools.partial(finish, j) for j in jobs]
        return [finish(j) for j in jobs]


def test_deferred_png_encoding_overlaps_the_next_gpu_pass(pool_factory):
    """the worker thread must not sit in PNG encoding.  Two incompatible
    (different size) requests = two GPU passes; with deferred results the second pass starts
    before the first one's encode has finished, and both futures resolve correctly."""
    w = DeferredWorker(0, encode_s=0.3)
    pool = pool_factory(worker_factory=lambda worker_id: w, num_workers=1, max_batch=16)
    t0 = time.perf_counter()
    f1 = pool.submit_job(GenerationJob(req=req("a", size="512x512")))
    f2 = pool.submit_job(GenerationJob(req=req("b", size="768x768")))
    assert f1.result(timeout=5) == (b"png-a", 0) and f2.result(timeout=5) == (b"png-b", 0)
    total = time.perf_counter() - t0
    assert len(w.gpu_done) == 2 and w.gpu_done[1] - w.gpu_done[0] < 0.25     # no wait for the encode
    assert total < 0.55                                                      # 2 x 0.3 s encodes overlapped
    # an exception raised on the encoder thread reaches the caller's future; the pool keeps serving
    fb = pool.submit_job(GenerationJob(req=req("boom")))
    with pytest.raises(ValueError, match="encode failed"):
        fb.result(timeout=5)
    assert pool.submit_job(GenerationJob(req=req("c"))).result(timeout=5) == (b"png-c", 0)


def test_cfg_workers_batch_by_guidance_scale(pool_factory):
    """A worker that runs classifier-free guidance on a doubled batch (B200SDXLWorker) needs one
    guidance scale per pass: the pool then splits the FIFO prefix where the scale changes."""
    class CfgWorker(FakeWorker):
        batch_same_guidance = True

    w = CfgWorker(0, delay=0.2)
    pool = pool_factory(worker_factory=lambda worker_id: w, num_workers=1, max_batch=16)
    blocker = pool.submit_job(GenerationJob(req=req("warm")))
    time.sleep(0.05)                                   # the worker is busy: the next six queue up
    futs = []
    for i, gs in enumerate([7.5, 7.5, 7.5, 5.0, 5.0, 7.5]):
        r = req(f"p{i}")
        r.guidance_scale = gs
        futs.append(pool.submit_job(GenerationJob(req=r)))
    blocker.result(timeout=5)
    assert [f.result(timeout=5)[0] for f in futs] == [f"png-p{i}".encode() for i in range(6)]
    assert w.batches == [1, 3, 2, 1]


def test_batch_window_collects_requests_below_saturation(pool_factory, monkeypatch):
    """Four idle workers, requests trickling in 3 ms apart: with the collector role + batch window the first
    idle thread gathers them into one batch instead of four threads taking one each."""
    monkeypatch.setenv("B200_BATCH_WINDOW_MS", "60")
    workers = {}

    def factory(worker_id):
        workers[worker_id] = FakeWorker(worker_id, delay=0.01)
        return workers[worker_id]
    pool = pool_factory(worker_factory=factory, num_workers=4, max_batch=8)
    assert abs(pool.batch_window - 0.060) < 1e-9
    futs = []
    for i in range(6):
        futs.append(pool.submit_job(GenerationJob(req=req(f"w{i}"))))
        time.sleep(0.003)
    assert [f.result(timeout=5)[0] for f in futs] == [f"png-w{i}".encode() for i in range(6)]
    sizes = sorted(b for w in workers.values() for b in w.batches)
    assert sizes == [6], sizes
    # window 0: never waits (the reference's latency for a lone request)
    monkeypatch.setenv("B200_BATCH_WINDOW_MS", "0")
    pool0 = pool_factory(worker_factory=lambda worker_id: FakeWorker(worker_id), num_workers=1, max_batch=8)
    t0 = time.perf_counter()
    assert pool0.submit_job(GenerationJob(req=req("solo"))).result(timeout=5)[0] == b"png-solo"
    assert time.perf_counter() - t0 < 0.5


def test_failed_mode_switch_keeps_the_old_workers(pool_factory):
    """A model that fails to load must not leave the pool without workers: the switch future
    carries the error, the old mode keeps serving, MODEL / MODEL_ROOT are restored."""
    import os
    calls = []

    def factory(worker_id):
        calls.append(os.environ.get("MODEL"))
        if os.environ.get("MODEL") == "alt":
            raise RuntimeError("b200 worker needs a diffusers-layout model directory")
        return FakeWorker(worker_id)
    pool = pool_factory(worker_factory=factory, num_workers=2, max_batch=4)
    old = list(pool._workers)
    with pytest.raises(RuntimeError, match="diffusers-layout"):
        pool.switch_mode("sd15-alt").result(timeout=5)
    assert pool.get_current_mode() == "sd15-fast" and pool._workers == old
    assert os.environ["MODEL"] == "sd15"
    assert pool.submit_job(GenerationJob(req=req("still"))).result(timeout=5)[0] == b"png-still"


def test_workers_are_created_concurrently(pool_factory):
    """8 GPU workers each load their own copy of the weights: serial construction would make start-up and
    every mode switch 8x longer."""
    def factory(worker_id):
        time.sleep(0.2)
        return FakeWorker(worker_id)
    t0 = time.perf_counter()
    pool = pool_factory(worker_factory=factory, num_workers=4, max_batch=4)
    assert time.perf_counter() - t0 < 0.6 and [w.worker_id for w in pool._workers] == [0, 1, 2, 3]


def test_deferred_results_with_a_wait_are_parked_on_one_waiter_per_batch(pool_factory):
    """A worker whose deferred results are still on the device hands out thunks with `wait()`: the pool parks ONE
    waiter thread per batch on it (not one encoder thread per request), the worker thread is back for the next batch
    at once, and every future still resolves; an error raised by the wait reaches every request of that batch."""
    gate = threading.Event()
    waits, order = [], []

    class Thunk:
        def __init__(self, tag, fail):
            self.tag, self.fail = tag, fail

        def wait(self):
            waits.append(self.tag)
            gate.wait(10)
            if self.fail:
                raise RuntimeError("device fault")

        def __call__(self):
            return (b"png-" + self.tag.encode(), 0)

    class W(FakeWorker):
        supports_deferred = True

        def run_batch(self, jobs, with_latents=False, deferred=False):
            order.append([j.req.prompt for j in jobs])
            fail = jobs[0].req.prompt.startswith("bad")
            th = [Thunk(j.req.prompt, fail) for j in jobs]
            return th if deferred else [t() for t in th]

    pool = pool_factory(worker_factory=lambda worker_id=0: W(worker_id), num_workers=1, max_batch=4)
    try:
        futs = [pool.submit_job(GenerationJob(req=req(f"a{i}"))) for i in range(4)]
        t0 = time.monotonic()
        while len(order) < 1 and time.monotonic() - t0 < 5:
            time.sleep(0.01)
        futs += [pool.submit_job(GenerationJob(req=req(f"bad{i}"))) for i in range(2)]
        t0 = time.monotonic()
        while sum(map(len, order)) < 6 and time.monotonic() - t0 < 5:   # later batches are taken while the first still "runs"
            time.sleep(0.01)
        assert len(order) >= 2 and not any(f.done() for f in futs)
        gate.set()
        assert sorted(f.result(timeout=10)[0] for f in futs[:4]) == [b"png-a0", b"png-a1", b"png-a2", b"png-a3"]
        for f in futs[4:]:
            with pytest.raises(RuntimeError, match="device fault"):
                f.result(timeout=10)
        assert len(waits) == len(order) < 6                        # one wait per batch, not per request
    finally:
        gate.set()
