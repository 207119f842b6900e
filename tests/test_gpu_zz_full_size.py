"""GPU parity at BASELINE config 5's domain size: the 1 003 000-cell synthetic domain of
`bench.py` (hourly forcing, routing case 1 on a 1 M-node network), one day of model steps, two
members.  The oracle still finishes this in seconds, so the comparison is direct; on top of it
the size-independent properties: members with the same parameters are bit-identical, and an
uneven split into time blocks equals a single block bit for bit."""
import numpy as np
import pytest

import orc_run
import parity
from mhm_b200 import driver, interface, synth

pytestmark = pytest.mark.gpu


def test_full_size_domain_equals_oracle():
    prob = synth.make_problem(nx=1180, ny=1000, n_days=1, hourly=True)
    assert prob["nCells"] > 1000000
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    assert np.abs(o.mRM_runoff).max() > 0
    members = [prob["params"], prob["params"]]
    with interface.Context() as ctx:
        ctx.set_math_mode("fast")
        dom = driver.setup_domain(ctx, 1, prob, nMembers=2, member_params=members)
        dom.run_steps(1, 7)
        dom.run_steps(8, nT - 7)
        q = [dom.get_runoff(member=m) for m in range(2)]
        worst = parity.assert_close(q[0], o.mRM_runoff, "gauge discharge, 1 M cells", rtol=parity.RTOL_Q)
        parity.assert_bit_exact(q[1], q[0], "gauge discharge member 1 vs member 0")
        S = {}
        for name in ("L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"):
            S[name] = dom.get_state(name, member=0)
            parity.assert_close(S[name], o.S[name], name + ", 1 M cells")
            parity.assert_bit_exact(dom.get_state(name, member=1), S[name], name + " member 1 vs member 0")
        R = {}
        for name in ("L11_qOUT", "L11_qTIN", "L11_qTR", "L11_qMod"):
            R[name] = dom.get_routing_state(name, member=0)
            parity.assert_close(R[name], o.R[name], name + ", 1 M nodes", rtol=parity.RTOL_Q)
        print("1 M cells x %d steps: gauge discharge max relative difference %.3e" % (nT, worst))
        interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, 1))
        del ctx.domains[1]
        dom = driver.setup_domain(ctx, 1, prob, nMembers=2, member_params=members)
        dom.run_steps(1, nT)
        parity.assert_bit_exact(dom.get_runoff(member=0), q[0], "gauge series, single block vs split")
        for name, v in S.items():
            parity.assert_bit_exact(dom.get_state(name, member=1), v, name + " single block vs split")
        for name, v in R.items():
            parity.assert_bit_exact(dom.get_routing_state(name, member=1), v, name + " single block vs split")
