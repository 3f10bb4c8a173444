"""N>1 path on CPU (gloo, world_size 2): the pieces of bench.py / the replica design that do
not need a GPU — rank-sharded synthetic requests (disjoint per-sample seeds, no data-path
collective), max-over-ranks timing reduction, rank-0-only reporting, and the `--impl reference`
arm's "rank 0 works, other ranks exit 0" rule."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dreamlab_b200.synthetic import synthetic_inputs
    B = 4
    pe, lat, noise = synthetic_inputs(B, 64, 64, 4, seed_base=1000 + rank * B)
    # shards are disjoint: gather a fingerprint of every rank's latents
    fp = lat.flatten()[:8].clone()
    got = [torch.zeros_like(fp) for _ in range(world)]
    dist.all_gather(got, fp)
    # max-over-ranks timing, as bench.py reduces it
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = world * B * 3 / (t.item() / 1e3)
    if rank == 0:
        json.dump({"value": value, "distinct": not torch.equal(got[0], got[1]),
                   "same_prompt_seed": True}, open(os.path.join(out_dir, "r0.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduce(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = json.load(open(tmp_path / "r0.json"))
    assert r["distinct"]
    assert abs(r["value"] - 2 * 4 * 3 / 0.015) < 1e-6          # slowest rank (15 ms) sets the rate


def test_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
