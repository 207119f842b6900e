"""Sub-catchment sharding on the device: N shards (one context each, on one GPU) exchanging
the cut links' outflow series per time block through an in-process stand-in for
torch.distributed; gauge series and node states must be bit-identical to the unsharded run.
The real NCCL exchange is exercised by `bench.py --shard` under torchrun."""
import numpy as np
import pytest
import torch

import parity
from mhm_b200 import driver, interface, shard, synth

pytestmark = pytest.mark.gpu


class FakeDist:
    """send/recv between shards that live in one process"""

    def __init__(self):
        self.box = {}
        self.cur = None

    def send(self, t, dst):
        self.box.setdefault(self.cur, []).append(t.clone())

    def recv(self, t, src):
        t.copy_(self.box[src].pop(0))


@pytest.mark.parametrize("n_parts,mode,members", [(2, "strict", 1), (4, "strict", 2), (3, "fast", 1)])
def test_sharded_domain_bit_identical_to_unsharded(n_parts, mode, members):
    prob = synth.make_problem(nx=50, ny=36, n_days=5, hourly=True)
    nT = prob["time"]["nTimeSteps"]
    rng = np.random.default_rng(7)
    mp = [prob["params"]] + [{k: (v * rng.uniform(0.9, 1.1) if k in ("L1_kPerco", "L1_kSlowFlow") else v)
                              for k, v in prob["params"].items()} for _ in range(members - 1)]
    with interface.Context() as ctx:
        ctx.set_math_mode(mode)
        dom = driver.setup_domain(ctx, 1, prob, nMembers=members, member_params=mp)
        dom.run_steps(1, nT)
        want_q = [dom.get_runoff(member=m) for m in range(members)]
        want_qmod = [dom.get_routing_state("L11_qMod", member=m) for m in range(members)]
        want_sm = dom.get_state("L1_soilMoist")
    part = shard.partition(prob["net"], n_parts)
    fd = FakeDist()
    ctxs = [interface.Context() for _ in range(n_parts)]
    try:
        runs = []
        for r, c in enumerate(ctxs):
            c.set_math_mode(mode)
            sub = shard.extract(prob, part, r)
            sub_mp = [{k: (np.ascontiguousarray(v[..., sub["shard"]["cells"]]) if k != "rout_param" else v)
                       for k, v in P.items()} for P in mp]
            run = shard.ShardedRun.__new__(shard.ShardedRun)
            run.torch, run.dist, run.rank, run.world, run.ctx = torch, fd, r, n_parts, c
            run.sub, run.M = sub, members
            run.dom = driver.setup_domain(c, 1, sub, nMembers=members, member_params=sub_mp)
            interface.check(c.L.mrm_cuda_set_deferred(c.h, 1, 1))
            run.device = torch.device("cuda", 0)
            runs.append(run)
        for first, n in ((1, 40), (41, 17), (58, nT - 57)):  # three time blocks
            for r in list(range(1, n_parts)) + [0]:
                fd.cur = r
                runs[r].run_block(first, n)
        for m in range(members):
            q = sum(np.where(np.isin(np.arange(want_q[m].shape[0])[:, None],
                                     np.asarray(run.sub["net"]["gaugeIndexList"]) - 1),
                             run.dom.get_runoff(member=m), 0.0) for run in runs)
            parity.assert_bit_exact(q, want_q[m], "gauge series, member %d" % m)
            qmod = np.zeros_like(want_qmod[m])
            for run in runs:
                nodes = run.sub["shard"]["nodes"]
                qmod[nodes] = run.dom.get_routing_state("L11_qMod", member=m)[: len(nodes)]
            parity.assert_bit_exact(qmod, want_qmod[m], "qMod of every node, member %d" % m)
        sm = np.zeros_like(want_sm)
        for run in runs:
            sm[:, run.sub["shard"]["cells"]] = run.dom.get_state("L1_soilMoist")
        parity.assert_bit_exact(sm, want_sm, "soil moisture")
        assert sum(len(np.asarray(r.sub["net"]["gaugeIndexList"])) for r in runs) == prob["net"]["nGaugesTotal"]
    finally:
        for c in ctxs:
            c.finalize()
