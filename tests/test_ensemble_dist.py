"""world_size-2 gloo test (CPU) of the multi-GPU host logic: member partition and the gather of
the members' gauge series on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mhm_b200 import ensemble


def test_partition_members():
    assert ensemble.partition_members(256, 8) == [(32 * r, 32) for r in range(8)]
    assert ensemble.partition_members(5, 2) == [(0, 3), (3, 2)]
    assert ensemble.partition_members(1, 4) == [(0, 1), (1, 0), (1, 0), (1, 0)]
    for n, w in ((7, 3), (64, 8), (3, 8)):
        p = ensemble.partition_members(n, w)
        assert sum(c for _, c in p) == n and all(p[i][0] + p[i][1] == p[i + 1][0] for i in range(w - 1))


def _worker(rank, world, port, n_members, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = ensemble.partition_members(n_members, world)[rank]
    # member m's series is m + gauge/10 + step/1000: recognisable after the gather
    g, s = np.meshgrid(np.arange(3), np.arange(20), indexing="ij")
    local = np.stack([m + g / 10.0 + s / 1000.0 for m in range(first, first + cnt)]) if cnt else np.zeros((0, 3, 20))
    out = ensemble.gather_runoff(local, n_members, dist)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_members", [5, 8])
def test_gather_runoff_gloo_world2(n_members):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_members, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out.shape == (n_members, 3, 20)
    g, s = np.meshgrid(np.arange(3), np.arange(20), indexing="ij")
    for m in range(n_members):
        assert np.array_equal(out[m], m + g / 10.0 + s / 1000.0)


def test_gather_single_process():
    x = np.random.default_rng(0).random((4, 2, 6))
    assert ensemble.gather_runoff(x, 4) is x
