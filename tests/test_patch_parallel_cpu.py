"""Host logic of the SDXL patch-parallel path on CPU: the strip communicator
primitives (halo exchange X1, row gather X2, record gather X3) over gloo world_size 2 and over
the in-process thread communicator, the rank topology (CFG halves x strips) and the assembly of
the per-step noise prediction — checked against a plain conv / slicing of the full tensor."""
import os
import sys
import threading

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_conv_check(comm, full, weight):
    """3x3 conv of `full` [B,C,H,W] computed strip-wise with exchanged halos == slice of the
    full conv."""
    B, C, H, W = full.shape
    hl = H // comm.world
    r0 = comm.rank * hl
    strip = full[:, :, r0:r0 + hl].permute(0, 2, 3, 1).contiguous()            # NHWC strip
    t = torch.empty(B, hl + 2, W, C)
    t[:, 1:hl + 1] = strip
    t[:, 0] = 7.0                                                                # garbage: must be overwritten
    t[:, hl + 1] = -7.0
    comm.halo_exchange(t)
    x = t.permute(0, 3, 1, 2)                                                    # [B,C,hl+2,W]
    y = F.conv2d(x, weight, padding=(0, 1))                                      # valid in H: halos supply the rows
    ref = F.conv2d(full, weight, padding=1)[:, :, r0:r0 + hl]
    assert torch.allclose(y, ref, atol=1e-5), (y - ref).abs().max()


def _gather_checks(comm):
    from dreamlab_b200.patch_parallel import Topology, assemble_eps
    B, n, Fd = 2, 3, 5
    full = torch.arange(B * comm.world * n * Fd, dtype=torch.float32).reshape(B, comm.world * n, Fd)
    mine = full[:, comm.rank * n:(comm.rank + 1) * n]
    assert torch.equal(comm.gather_rows(mine), full)
    assert torch.equal(comm.gather_rows(mine[:1]), full[:1])                    # B == 1: view path
    rec = torch.full((B, 4, 2), float(comm.rank))
    g = comm.all_gather(rec)
    assert g.shape == (comm.world, B, 4, 2) and all(float(g[r].mean()) == r for r in range(comm.world))
    # no CFG: world of strips
    topo = Topology(comm.world, comm.rank, cfg=False)
    assert (topo.cfg_ways, topo.strips, topo.strip_index) == (1, comm.world, comm.rank)
    eps_full = torch.randn(B, comm.world * 2, 4, 4, generator=torch.Generator().manual_seed(3))
    strip = eps_full[:, comm.rank * 2:(comm.rank + 1) * 2]
    e, none = assemble_eps(comm.all_gather(strip), topo, B)
    assert none is None and torch.equal(e, eps_full)


def _gloo_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dreamlab_b200.patch_parallel import DistComm, Topology, assemble_eps
    comm = DistComm()
    assert (comm.rank, comm.world) == (rank, world)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(2, 8, 8, 6, generator=g)
    weight = torch.randn(4, 8, 3, 3, generator=g)
    _strip_conv_check(comm, full, weight)
    _gather_checks(comm)
    # CFG split on 2 ranks: rank 0 = unconditional half, rank 1 = conditional half, 1 strip each
    topo = Topology(world, rank, cfg=True)
    assert (topo.cfg_ways, topo.strips, topo.cfg_index, topo.strip_index) == (2, 1, rank, 0)
    half = torch.full((1, 4, 4, 4), float(rank))
    e_u, e_t = assemble_eps(comm.all_gather(half), topo, 1)
    assert float(e_u.mean()) == 0.0 and float(e_t.mean()) == 1.0 and e_u.shape == (1, 4, 4, 4)
    dist.barrier()
    dist.destroy_process_group()


def test_strip_comm_over_gloo_world2():
    port = 31500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port), nprocs=2, join=True)


def test_thread_comm_world4_with_cfg_halves():
    sys.path.insert(0, ROOT)
    from dreamlab_b200.patch_parallel import ThreadComm, Topology, _ThreadHub, assemble_eps
    world = 4
    comms = ThreadComm.make(world)
    hubs = [_ThreadHub(2), _ThreadHub(2)]
    g = torch.Generator().manual_seed(0)
    full = torch.randn(1, 8, 16, 6, generator=g)
    weight = torch.randn(4, 8, 3, 3, generator=g)
    eps_u = torch.randn(1, 8, 4, 4, generator=g)
    eps_t = torch.randn(1, 8, 4, 4, generator=g)
    errs = []

    def run(rank):
        try:
            comm = comms[rank]
            _strip_conv_check(comm, full, weight)
            _gather_checks(comm)
            topo = Topology(world, rank, cfg=True)
            assert (topo.cfg_ways, topo.strips) == (2, 2)
            assert topo.strip_ranks(topo.cfg_index) == ([0, 1] if rank < 2 else [2, 3])
            sub = ThreadComm(hubs[topo.cfg_index], topo.strip_index)
            _strip_conv_check(sub, full[:, :, :8], weight)        # halos stay inside the half's group
            src = eps_u if topo.cfg_index == 0 else eps_t
            strip = src[:, topo.strip_index * 4:(topo.strip_index + 1) * 4]
            a, b = assemble_eps(comm.all_gather(strip), topo, 1)
            assert torch.equal(a, eps_u) and torch.equal(b, eps_t)
        except BaseException as e:           # noqa: BLE001 - surface in the main thread
            errs.append((rank, repr(e)))
            for c in comms:
                c.hub.barrier.abort()
            for h in hubs:
                h.barrier.abort()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errs, errs


def test_single_comm_halo_is_zero_padding():
    sys.path.insert(0, ROOT)
    from dreamlab_b200.patch_parallel import SingleComm
    g = torch.Generator().manual_seed(0)
    _strip_conv_check(SingleComm(), torch.randn(1, 8, 6, 6, generator=g), torch.randn(4, 8, 3, 3, generator=g))
