"""CPU tests of the oracle's time loop: restartability and routing invariants."""
import numpy as np

import orc_run
import parity
from mhm_b200 import synth


def test_oracle_split_run_is_identical():
    prob = synth.make_problem(nx=12, ny=8, n_days=5, hourly=False, rout_case=1, l1_factor=2)
    nT = prob["time"]["nTimeSteps"]
    a = orc_run.OracleRun(prob)
    a.run(1, nT)
    b = orc_run.OracleRun(prob)
    b.run(1, 37)
    b.run(38, nT)
    parity.assert_bit_exact(a.mRM_runoff, b.mRM_runoff, "gauge series")
    for k in a.S:
        parity.assert_bit_exact(a.S[k], b.S[k], k)
    assert a.mRM_runoff.max() > 0


def test_oracle_threads_do_not_change_results():
    prob = synth.make_problem(nx=12, ny=8, n_days=3, hourly=True)
    nT = prob["time"]["nTimeSteps"]
    a = orc_run.OracleRun(prob, num_threads=1)
    a.run(1, nT)
    b = orc_run.OracleRun(prob, num_threads=4)
    b.run(1, nT)
    parity.assert_bit_exact(a.mRM_runoff, b.mRM_runoff, "gauge series")


def test_water_balance_closes():
    """size-independent property: P = ET + runoff + storage change (per cell, whole run)."""
    prob = synth.make_problem(nx=10, ny=8, n_days=8, hourly=True, routing=False)
    # karstLoss < 1 removes water from the balance; make it conservative
    prob["params"]["L1_karstLoss"][:] = 1.0
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob, history=True)
    o.run(1, nT)
    y = [prob["time"]["LCyearId"][0] - 1]
    P = prob["params"]
    fs_by_step = np.stack([P["L1_fSealed"][yy - 1, 0] for yy in o.time_indices(nT)["yId"]])
    H = lambda name: np.stack([o.hist(name, tt) for tt in range(1, nT + 1)])
    prec = H("L1_prec_calc").sum(0)
    et = (H("L1_aETCanopy") + H("L1_aETSealed") * fs_by_step
          + H("L1_aETSoil").sum(1) * (1 - fs_by_step)).sum(0)
    ro = H("L1_total_runoff").sum(0)
    n = prob["nCells"]

    def storage(tt, fs):
        if tt == 0:
            s0 = prob["states0"]
            sm = 0.5 * P["L1_soilMoistFC"][y[0]]
            return (s0["L1_inter"] + s0["L1_snowPack"] + s0["L1_sealSTW"] * fs
                    + (sm.sum(0) + s0["L1_unsatSTW"] + s0["L1_satSTW"]) * (1 - fs))
        return (o.hist("L1_inter", tt) + o.hist("L1_snowPack", tt) + o.hist("L1_sealSTW", tt) * fs
                + (o.hist("L1_soilMoist", tt).sum(0) + o.hist("L1_unsatSTW", tt)
                   + o.hist("L1_satSTW", tt)) * (1 - fs))

    # the sealed fraction changes with the land-cover scene; restrict to cells where it is equal
    same = np.isclose(P["L1_fSealed"][0, 0], P["L1_fSealed"][-1, 0])
    fs = P["L1_fSealed"][0, 0]
    ds = storage(nT, fs) - storage(0, fs)
    resid = prec - et - ro - ds
    if same.any():
        assert np.abs(resid[same]).max() < 1e-8
    # all cells with fSealed == 0 in both scenes must close exactly as well
    unsealed = (P["L1_fSealed"][:, 0] == 0).all(0)
    assert unsealed.any() and np.abs(resid[unsealed]).max() < 1e-8
