"""Sub-catchment sharding of ONE domain under real NCCL, exchange inside the library
(mrm_cuda_set_exchange / mrm_cuda_shard_run_steps): called by tests/multi_gpu_worker.py, one rank
per GPU.  Every rank also runs the whole domain unsharded on its own GPU; what the shard owns
(gauge columns, node states, cell states) must equal that run bit for bit."""
import numpy as np

import parity
from mhm_b200 import driver, interface, shard, synth


def run(ctx, rank, world, dist):
    for mode, members in (("strict", 1), ("fast", 2)):
        prob = synth.make_problem(nx=64, ny=44, n_days=6, hourly=True, n_gauges=6)
        nT = prob["time"]["nTimeSteps"]
        rng = np.random.default_rng(7)
        mp = [prob["params"]] + [{k: (v * rng.uniform(0.9, 1.1) if k in ("L1_kPerco", "L1_kSlowFlow") else v)
                                  for k, v in prob["params"].items()} for _ in range(members - 1)]
        ctx.set_math_mode(mode)
        dom = driver.setup_domain(ctx, 7, prob, nMembers=members, member_params=mp)
        dom.run_steps(1, nT)
        want_q = [dom.get_runoff(member=m) for m in range(members)]
        want_qmod = [dom.get_routing_state("L11_qMod", member=m) for m in range(members)]
        want_sm = [dom.get_state("L1_soilMoist", member=m) for m in range(members)]
        interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, 7))
        del ctx.domains[7]
        part = shard.partition(prob["net"], world)
        sr = shard.ShardedRun(ctx, prob, part, rank, world, dist, nMembers=members, member_params=mp)
        assert sr.native
        for first, n in ((1, 40), (41, 17), (58, nT - 57)):  # three uneven time blocks
            sr.run_block(first, n)
        sr.finish()
        ctx.synchronize()
        sh = sr.sub["shard"]
        cols = np.asarray(sr.sub["net"]["gaugeIndexList"], dtype=np.int64) - 1
        for m in range(members):
            q = sr.dom.get_runoff(member=m)
            parity.assert_bit_exact(q[cols], want_q[m][cols], "shard %d: gauge series, member %d" % (rank, m))
            qmod = sr.dom.get_routing_state("L11_qMod", member=m)[: len(sh["nodes"])]
            parity.assert_bit_exact(qmod, want_qmod[m][sh["nodes"]], "shard %d: qMod, member %d" % (rank, m))
            sm = sr.dom.get_state("L1_soilMoist", member=m)
            parity.assert_bit_exact(sm, np.ascontiguousarray(want_sm[m][:, sh["cells"]]), "shard %d: soil moisture" % rank)
        # every gauge has exactly one owner
        owned = np.zeros(prob["net"]["nGaugesTotal"], dtype=np.int64)
        owned[cols] = 1
        t = __import__("torch").from_numpy(owned).cuda()
        dist.all_reduce(t)
        assert (t.cpu().numpy() == 1).all()
        interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, 1))
        del ctx.domains[1]
        if rank == 0:
            print("MULTI_GPU_OK sharded domain (%s, %d members): %d shards, %d cut links, exchange inside the library" % (
                mode, members, world, sh["n_ghost"]), flush=True)
