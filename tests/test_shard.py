"""Sub-catchment sharding: the partitioner, the sub-network extraction and the
exchange protocol (one message per shard and time block with the outflow series of the cut
links), tested on CPU: world_size-2 gloo run of a ghost-aware numpy Muskingum sweep against the
same sweep on the whole network (bit-identical), which itself is checked against the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc_run
from mhm_b200 import shard, synth
from mhm_b200.interface import routing_order


def muskingum_sweep(net, C1, C2, qout, ghost=None, last_sink=0):
    """L11_routing (mRM/mo_mrm_routing.f90:428-478) over all steps of qout[t][node]; ghost =
    {local node (0-based): outflow series} replaces the routed outflow of ghost sources.
    Returns qTR series per node (of its outgoing link) and qMod = qTIN(:, 2) per step."""
    nn = net["nNodes"]
    nl = nn - net["nOutlets"]
    T = qout.shape[0]
    tin1, tr1 = np.zeros(nn), np.zeros(nn)
    qtr, qmod = np.zeros((T, nn)), np.zeros((T, nn))
    ghost = ghost or {}
    last = None
    for t in range(T):
        tin2, tr2 = np.zeros(nn), np.zeros(nn)
        for k in range(nl):
            i = net["netPerm"][k] - 1
            f, to = net["fromN"][i] - 1, net["toN"][i] - 1
            if f in ghost:
                tr2[f] = ghost[f][t]
            else:
                tin2[f] = tin2[f] + qout[t, f]
                tr2[f] = tr1[f] + C1[i] * (tin1[f] - tr1[f]) + C2[i] * (tin2[f] - tin1[f])
            tin2[to] = tin2[to] + tr2[f]
            last = to
        if last_sink > 0:
            last = last_sink - 1
        if nl > 0 and last_sink >= 0:
            tin2[last] = tin2[last] + qout[t, last]
        qtr[t], qmod[t] = tr2, tin2
        tin1, tr1 = tin2, tr2
    return qtr, qmod


def make(nx=36, ny=24, T=30, seed=5):
    prob = synth.make_problem(nx=nx, ny=ny, n_days=2, hourly=True, routing_order=routing_order, seed=seed)
    net = prob["net"]
    rng = np.random.default_rng(seed)
    nl = net["nNodes"] - net["nOutlets"]
    C1, C2 = rng.uniform(0.2, 1.0, net["nNodes"]), rng.uniform(0.0, 0.5, net["nNodes"])
    qout = rng.gamma(0.7, 2.0, (T, net["nNodes"]))
    return prob, C1, C2, qout, nl


@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_partition_covers_and_cuts_end_in_trunk(n_parts):
    prob, _, _, _, nl = make(nx=80, ny=50)
    net = prob["net"]
    part = shard.partition(net, n_parts)
    assert part.min() == 0 and part.max() == n_parts - 1
    load = np.bincount(part, minlength=n_parts)
    assert load.max() <= 1.15 * load.mean(), load
    f, t = net["fromN"][:nl] - 1, net["toN"][:nl] - 1
    cut = part[f] != part[t]
    assert (part[t[cut]] == 0).all() and cut.sum() > 0
    subs = [shard.extract(prob, part, r) for r in range(n_parts)]
    assert sum(s["nCells"] for s in subs) == prob["nCells"]
    assert subs[0]["shard"]["n_ghost"] == cut.sum() == sum(s["shard"]["n_export"] for s in subs)
    for s in subs:  # local netPerm is a topological order of the local links
        ln = s["net"]
        nll = ln["nNodes"] - ln["nOutlets"]
        seen = np.zeros(ln["nNodes"], dtype=bool)
        has_up = np.zeros(ln["nNodes"], dtype=bool)
        has_up[ln["toN"][:nll] - 1] = True
        for k in range(nll):
            i = ln["netPerm"][k] - 1
            seen[ln["fromN"][i] - 1] = True
        assert seen[ln["fromN"][:nll] - 1].all()


def sharded_sweep(prob, part, C1, C2, qout, rank, send, recv):
    """route shard `rank`; send(buf) / recv(src_rank, n) move cut-link series between shards"""
    sub = shard.extract(prob, part, rank)
    ln, sh = sub["net"], sub["shard"]
    nl = prob["net"]["nNodes"] - prob["net"]["nOutlets"]
    # link-indexed C1/C2 and node-indexed runoff of the local network
    f_glob = prob["net"]["fromN"][:nl] - 1
    own = part == rank
    cutm = part[f_glob] != part[prob["net"]["toN"][:nl] - 1]
    links_own = np.nonzero(own[f_glob])[0]
    ghost_links = np.nonzero(cutm & ~own[f_glob])[0] if rank == 0 else np.zeros(0, np.int64)
    links = np.sort(np.concatenate([links_own, ghost_links])).astype(np.int64)
    c1 = np.zeros(ln["nNodes"])
    c2 = np.zeros(ln["nNodes"])
    c1[: len(links)], c2[: len(links)] = C1[links], C2[links]
    q = np.zeros((qout.shape[0], ln["nNodes"]))
    q[:, : len(sh["nodes"])] = qout[:, sh["nodes"]]
    ghost = {}
    if rank == 0:
        off = 0
        for r in range(1, sh["n_parts"]):
            c = sh["recv_counts"][r]
            if c:
                buf = recv(r, c)
                for j in range(c):
                    ghost[int(ln["ghostSourceNodeList"][off + j]) - 1] = buf[j]
                off += c
    qtr, qmod = muskingum_sweep(ln, c1, c2, q, ghost, ln["lastSinkNode"])
    if rank != 0 and sh["n_export"]:
        send(np.stack([qtr[:, int(e) - 1] for e in ln["exportNodeList"]]))
    out = np.zeros((qout.shape[0], prob["net"]["nNodes"]))
    out[:, sh["nodes"]] = qmod[:, : len(sh["nodes"])]
    return out


def test_numpy_sweep_equals_oracle_routing():
    """the helper used as the truth of the sharding tests is itself the oracle's routing"""
    prob = synth.make_problem(nx=20, ny=12, n_days=2, hourly=True, routing_order=routing_order)
    o = orc_run.OracleRun(prob, history=True)
    nT = prob["time"]["nTimeSteps"]
    o.run(1, nT)
    net = prob["net"]
    runoff = np.stack([o.hist("L1_total_runoff", tt) for tt in range(1, nT + 1)])
    qout = (0.0 + runoff * net["L1_areaCell"][None, :]) * 1000.0 / 3600.0
    _, qmod = muskingum_sweep(net, o.R["L11_C1"], o.R["L11_C2"], qout)
    g = net["gaugeNodeList"] - 1
    assert np.array_equal(qmod[:, g].T, o.mRM_runoff)


@pytest.mark.parametrize("n_parts", [2, 4])
def test_sharded_sweep_in_process_bit_identical(n_parts):
    prob, C1, C2, qout, _ = make()
    part = shard.partition(prob["net"], n_parts)
    _, want = muskingum_sweep(prob["net"], C1, C2, qout)
    box = {}
    got = np.zeros_like(want)
    for rank in list(range(1, n_parts)) + [0]:
        got += sharded_sweep(prob, part, C1, C2, qout, rank, lambda b, r=rank: box.__setitem__(r, b),
                             lambda src, n: box[src])
    assert np.array_equal(got, want)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob, C1, C2, qout, _ = make()
    part = shard.partition(prob["net"], world)
    T = qout.shape[0]

    def send(buf):
        dist.send(torch.from_numpy(np.ascontiguousarray(buf)), dst=0)

    def recv(src, n):
        t = torch.empty((n, T), dtype=torch.float64)
        dist.recv(t, src=src)
        return t.numpy()

    mine = sharded_sweep(prob, part, C1, C2, qout, rank, send, recv)
    t = torch.from_numpy(mine)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)  # every node is owned by exactly one shard
    if rank == 0:
        q.put(t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sweep_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    prob, C1, C2, qout, _ = make()
    _, want = muskingum_sweep(prob["net"], C1, C2, qout)
    assert np.array_equal(got, want)
