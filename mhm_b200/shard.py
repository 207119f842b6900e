"""Sub-catchment sharding of ONE large domain over several GPUs.

L1 cells have no lateral coupling before routing, so the cell kernel and MPR shard by cell
without any exchange.  The river network is cut at links: every shard owns whole sub-catchments
(`mrm_partition_subcatchments`), shard 0 additionally the trunk below all cuts.  The routed
outflow qTR(iNode, t) of every cut link (mRM/mo_mrm_routing.f90:443-457) is the only datum that
crosses shards.  Because the routing runs over a whole time block, the series of ALL steps of
the block is sent in one message per shard and block: shards 1..N-1 route, send, shard 0
receives, routes.  On shard 0 the from-node of a cut link is a *ghost source* whose link keeps
its place in netPerm, so the inflows of its to-node are summed in the reference's order and the
sharded run is bit-identical to the unsharded one.

Host-side helpers (numpy); the exchange itself is `torch.distributed` send/recv over NCCL
(`ShardedRun`), or plain device copies when the shards are domains of one context (tests).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .interface import _pi

CELL_KEYS_SKIP = ("rout_param",)


def partition(net, n_parts):
    """0-based shard id of every L11 node"""
    L = _lib.load()
    nn = net["nNodes"]
    nl = nn - net["nOutlets"]
    part = np.zeros(nn, dtype=np.int32)
    f = np.ascontiguousarray(net["fromN"], dtype=np.int32)
    t = np.ascontiguousarray(net["toN"], dtype=np.int32)
    p = np.ascontiguousarray(net["netPerm"], dtype=np.int32)
    L.mrm_partition_subcatchments.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]
    check(L.mrm_partition_subcatchments(nn, nl, _pi(f), _pi(t), _pi(p), n_parts, _pi(part)))
    return part


def extract(prob, part, rank):
    """the sub-problem of shard `rank` (same dict layout as mhm_b200.synth.make_problem) plus
    prob['shard'] = exchange metadata.  Needs L11 == L1 with one cell per node (map_flag)."""
    net = prob["net"]
    nn, nl = net["nNodes"], net["nNodes"] - net["nOutlets"]
    fromN, toN = np.asarray(net["fromN"][:nl]) - 1, np.asarray(net["toN"][:nl]) - 1
    perm = np.asarray(net["netPerm"][:nl]) - 1
    node_of_cell = np.asarray(net["L1_L11_Id"]) - 1
    assert net["map_flag"] and len(node_of_cell) == nn and len(np.unique(node_of_cell)) == nn
    own = part == rank
    cut = part[fromN] != part[toN]                      # cut links, global link order
    assert (part[toN[cut]] == 0).all(), "cut links must end in the trunk (shard 0)"
    owned_nodes = np.nonzero(own)[0]
    loc = np.full(nn, -1, dtype=np.int64)
    loc[owned_nodes] = np.arange(len(owned_nodes))
    n_local = len(owned_nodes)
    links_own = np.nonzero(own[fromN])[0]               # links whose from-node is owned
    ghost_links = np.nonzero(cut & (part[toN] == rank) & ~own[fromN])[0] if rank == 0 else np.zeros(0, np.int64)
    # ghosts grouped by source shard, then by global link index: the order of the receive buffer
    ghost_links = ghost_links[np.lexsort((ghost_links, part[fromN[ghost_links]]))]
    export_links = np.nonzero(cut & own[fromN])[0]      # ascending global link index
    ghost_id = {int(g): n_local + i for i, g in enumerate(ghost_links)}
    sink_id = {int(g): n_local + len(ghost_links) + i for i, g in enumerate(export_links)}
    nn_loc = n_local + len(ghost_links) + len(export_links)
    links = np.sort(np.concatenate([links_own, ghost_links])).astype(np.int64)
    lid = {int(g): i for i, g in enumerate(links)}
    f_loc = np.full(nn_loc, -9999, dtype=np.int32)
    t_loc = np.full(nn_loc, -9999, dtype=np.int32)
    for i, g in enumerate(links):
        g = int(g)
        f_loc[i] = (ghost_id[g] if g in ghost_id else loc[fromN[g]]) + 1
        t_loc[i] = (sink_id[g] if g in sink_id else loc[toN[g]]) + 1
    in_local = np.zeros(nl, dtype=bool)
    in_local[links] = True
    p_loc = np.full(nn_loc, -9999, dtype=np.int32)
    sel = perm[in_local[perm]]                          # global netPerm restricted to local links
    p_loc[: len(sel)] = np.array([lid[int(g)] for g in sel], dtype=np.int32) + 1
    cells = np.nonzero(own[node_of_cell])[0]            # ascending cell index

    def per_link(a):                                    # link-indexed (.., nNodes) arrays
        a = np.asarray(a)
        out = np.zeros(a.shape[:-1] + (nn_loc,), dtype=a.dtype)
        out[..., : len(links)] = a[..., links]
        return out

    sub = dict(prob)
    sub["nCells"] = len(cells)
    sub["params"] = {k: (np.ascontiguousarray(v[..., cells]) if k not in CELL_KEYS_SKIP else v)
                     for k, v in prob["params"].items()}
    sub["states0"] = {k: np.ascontiguousarray(v[..., cells]) for k, v in prob["states0"].items()}
    sub["forcing"] = {k: np.ascontiguousarray(v[:, cells]) for k, v in prob["forcing"].items()}
    if prob.get("read_weights"):
        sub["weights"] = {k: np.ascontiguousarray(v[..., cells]) for k, v in prob["weights"].items()}
    gsel = [g for g, nd in enumerate(np.asarray(net["gaugeNodeList"]) - 1) if own[nd]]
    area11 = np.zeros(nn_loc)
    area11[:n_local] = np.asarray(net["L11_areaCell"])[owned_nodes]
    # reg_rout's maxval(slope(:)) runs over L11_slope(s11:e11-1), i.e. nNodes - 1 entries of the WHOLE
    # domain -- with several outlets that is more than its links (mRM/mo_mrm_mpr.f90:97,
    # mHM/mo_mhm_interface_run.f90:577)
    sl = np.asarray(net["L11_slope"])[: max(nn - 1, 0)]
    lnet = {
        "nNodes": nn_loc, "nOutlets": nn_loc - len(links), "nCells1": len(cells), "map_flag": 1,
        "fromN": f_loc, "toN": t_loc, "netPerm": p_loc,
        "L1_L11_Id": (loc[node_of_cell[cells]] + 1).astype(np.int32),
        "L11_L1_Id": np.ones(nn_loc, dtype=np.int32),
        "L1_areaCell": np.ascontiguousarray(np.asarray(net["L1_areaCell"])[cells]), "L11_areaCell": area11,
        "gaugeNodeList": (loc[np.asarray(net["gaugeNodeList"])[gsel] - 1] + 1).astype(np.int32),
        "gaugeIndexList": np.asarray(net["gaugeIndexList"])[gsel].astype(np.int32),
        "nGaugesTotal": net["nGaugesTotal"],
        "InflowGaugeNodeList": np.zeros(0, np.int32), "InflowGaugeIndexList": np.zeros(0, np.int32),
        "InflowGaugeHeadwater": np.zeros(0, np.int32), "nInflowTotal": 0,
        "processCase": net["processCase"], "rout_param": net["rout_param"],
        "L11_length": per_link(net["L11_length"]), "L11_slope": per_link(net["L11_slope"]),
        "L11_nLinkFracFPimp": per_link(net["L11_nLinkFracFPimp"]),
        "ghostSourceNodeList": np.array([ghost_id[int(g)] + 1 for g in ghost_links], dtype=np.int32),
        "exportNodeList": np.array([loc[fromN[int(g)]] + 1 for g in export_links], dtype=np.int32),
        "ssMax": float(sl.max()) if len(sl) else 0.0,
    }
    # only the whole domain's last sink adds its own runoff (mo_mrm_routing.f90:466-467)
    g_last = int(toN[perm[nl - 1]]) if nl > 0 else -1
    lnet["lastSinkNode"] = int(loc[g_last]) + 1 if g_last >= 0 and own[g_last] else -1
    assert net["nInflowTotal"] == 0, "inflow gauges are not supported on sharded domains"
    for k in ("C1", "C2"):
        if k in net:
            lnet[k] = per_link(net[k])
    for k in ("TSrout", "celerity"):
        if k in net:
            lnet[k] = net[k]
    sub["net"] = lnet
    sub["inflowQ"] = np.zeros((0, prob["inflowQ"].shape[1]))
    # what shard 0 receives from every other shard: number of cut links per source shard
    n_parts = int(part.max()) + 1
    counts = [int((cut & (part[fromN] == r)).sum()) for r in range(n_parts)]
    sub["shard"] = {"rank": rank, "n_parts": n_parts, "cells": cells, "nodes": owned_nodes,
                    "n_export": len(export_links), "n_ghost": len(ghost_links),
                    "recv_counts": counts, "gauges": gsel}
    return sub


class ShardedRun:
    """one shard per rank; `dist` = torch.distributed (NCCL on GPUs) or None with world size 1"""

    def __init__(self, ctx, prob, part, rank, world, dist=None, device=None, nMembers=1, member_params=None,
                 native=None):
        """native: the exchange runs inside the library (mrm_cuda_set_exchange / mrm_cuda_shard_run_steps,
        NCCL on the library's stream, shard 0 one block behind); default when the context owns a
        communicator of `world` ranks.  Otherwise torch.distributed send / recv from Python."""
        import torch

        from . import driver

        self.torch, self.dist, self.rank, self.world, self.ctx = torch, dist, rank, world, ctx
        self.sub = extract(prob, part, rank)
        cells = self.sub["shard"]["cells"]
        sub_mp = None
        if member_params is not None:
            sub_mp = [{k: (np.ascontiguousarray(v[..., cells]) if k not in CELL_KEYS_SKIP else v)
                       for k, v in P.items()} for P in member_params]
        self.dom = driver.setup_domain(ctx, 1, self.sub, nMembers=nMembers, member_params=sub_mp)
        if native is None:
            native = world > 1 and ctx.comm_info()["nranks"] == world
        self.native = native
        if native:
            sh = self.sub["shard"]
            send = np.zeros(world, dtype=np.int32)
            recv = np.zeros(world, dtype=np.int32)
            if rank == 0:
                recv[: len(sh["recv_counts"])] = sh["recv_counts"]
                recv[0] = 0
            else:
                send[0] = sh["n_export"]
            check(ctx.L.mrm_cuda_set_exchange(ctx.h, 1, _pi(send), _pi(recv)))
        else:
            check(ctx.L.mrm_cuda_set_deferred(ctx.h, 1, 1))
        self.M = nMembers
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def finish(self):
        """route what is still pending (native exchange: shard 0 works one block behind)"""
        if getattr(self, "native", False):
            check(self.ctx.L.mrm_cuda_shard_flush(self.ctx.h, 1))

    def run_block(self, tt_first, n_steps):
        """cells of the block everywhere, then routing: shards 1.. first, shard 0 after the
        outflow series of all cut links have arrived"""
        t, L, ctx, sh = self.torch, self.ctx.L, self.ctx, self.sub["shard"]
        if getattr(self, "native", False):
            check(L.mrm_cuda_shard_run_steps(ctx.h, 1, tt_first, n_steps))
            return
        self.dom.run_steps(tt_first, n_steps)
        if self.rank != 0:
            check(L.mrm_cuda_route_pending(ctx.h, 1))
            if sh["n_export"]:
                buf = t.empty((self.M, sh["n_export"], n_steps), dtype=t.float64, device=self.device)
                check(L.mrm_cuda_export_outflow(ctx.h, 1, C.c_void_p(buf.data_ptr()), n_steps))
                ctx.synchronize()
                self.dist.send(buf, dst=0)
        else:
            if sh["n_ghost"]:
                full = t.empty((self.M, sh["n_ghost"], n_steps), dtype=t.float64, device=self.device)
                off = 0
                for r in range(1, self.world):
                    c = sh["recv_counts"][r]
                    if c == 0:
                        continue
                    buf = t.empty((self.M, c, n_steps), dtype=t.float64, device=self.device)
                    self.dist.recv(buf, src=r)
                    full[:, off:off + c, :] = buf
                    off += c
                t.cuda.synchronize()
                check(L.mrm_cuda_import_outflow(ctx.h, 1, C.c_void_p(full.data_ptr()), n_steps))
            check(L.mrm_cuda_route_pending(ctx.h, 1))

    def gauge_series(self):
        """this shard's columns of mRM_runoff (others zero): sum over shards = full series"""
        q = self.dom.get_runoff()
        keep = np.zeros_like(q)
        cols = np.asarray(self.sub["net"]["gaugeIndexList"], dtype=np.int64) - 1
        keep[cols] = q[cols]
        return keep
