"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Run on the B200 box: `pytest tests -m gpu`."""
import numpy as np
import pytest

import orc_run
import parity
from mhm_b200 import driver, interface, synth

pytestmark = pytest.mark.gpu

STATES = ["L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"]
FLUXES = orc_run.FLUX_ORDER + ["L1_aETSoil", "L1_infilSoil"]


@pytest.fixture(scope="module")
def ctx():
    c = interface.Context()
    yield c
    c.finalize()


def fresh(ctx, prob, **kw):
    for k in list(ctx.domains):
        interface.check(ctx.L.mhm_cuda_unregister_domain(ctx.h, k))
        del ctx.domains[k]
    return driver.setup_domain(ctx, 1, prob, **kw)


CELL_CASES = [
    # (hourly, soil_case, pet_case, nH, timestep_h, weights)
    (True, 1, -1, 2, 1, False),
    (True, 2, 0, 3, 1, False),
    (False, 1, 0, 2, 1, False),
    (False, 4, -1, 1, 1, False),
    (False, 3, 1, 2, 1, False),
    (False, 1, 2, 2, 1, False),
    (False, 2, 3, 2, 2, False),
    (False, 1, -1, 2, 1, True),
    (False, 1, 0, 4, 24, False),
]


@pytest.mark.parametrize("hourly,soil_case,pet_case,nH,timestep_h,weights", CELL_CASES)
def test_per_step_seam_all_fluxes_and_states(ctx, hourly, soil_case, pet_case, nH, timestep_h, weights):
    """B1: every flux and state after every step, strict mode, vs the oracle's history."""
    n_days = 4 if timestep_h == 1 else 12
    prob = synth.make_problem(nx=16, ny=10, n_days=n_days, nH=nH, hourly=hourly, soil_case=soil_case,
                              pet_case=pet_case, timestep_h=timestep_h, routing=False,
                              read_weights=weights)
    o = orc_run.OracleRun(prob, history=True)
    nT = prob["time"]["nTimeSteps"]
    o.run(1, nT)
    ctx.set_math_mode("strict")
    dom = fresh(ctx, prob)
    idx = interface.time_indices(prob["time"], timestep_h, prob["nTstepForcingDay"], 1, nT)
    worst = 0.0
    for tt in range(1, nT + 1):
        dom.do_time_step(tt, idx[tt - 1])
        for name in STATES + FLUXES:
            got = dom.get_variable(name)
            worst = max(worst, parity.assert_close(got, o.hist(name, tt), "%s @tt=%d" % (name, tt)))
    print("max relative difference %.3e" % worst)
    assert len({s.yId for s in idx}) == 2, "the run must cross a land-cover scene change"


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("hourly,soil_case,pet_case,nH", [(True, 1, -1, 2), (False, 2, 0, 3)])
def test_time_block_equals_oracle(ctx, mode, hourly, soil_case, pet_case, nH):
    """B2: a block of steps in one launch: total-runoff history of every step, final states
    and the last step's fluxes; also split blocks == one block (bit-exact)."""
    prob = synth.make_problem(nx=30, ny=20, n_days=10, nH=nH, hourly=hourly, soil_case=soil_case,
                              pet_case=pet_case, routing=False)
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob, history=True)
    o.run(1, nT)
    ctx.set_math_mode(mode)
    dom = fresh(ctx, prob)
    dom.run_steps(1, nT)
    hist = dom.get_runoff_history(nT)
    ref = np.stack([o.hist("L1_total_runoff", tt) for tt in range(1, nT + 1)])
    worst = parity.assert_close(hist, ref, "total_runoff history (%s)" % mode)
    final = {}
    for name in STATES + FLUXES:
        final[name] = dom.get_variable(name)
        worst = max(worst, parity.assert_close(final[name], o.hist(name, nT), name))
    print("%s: max relative difference %.3e" % (mode, worst))
    # same run in three uneven blocks
    dom = fresh(ctx, prob)
    dom.run_steps(1, 7)
    dom.run_steps(8, 100)
    dom.run_steps(108, nT - 107)
    for name in STATES + FLUXES:
        parity.assert_bit_exact(dom.get_variable(name), final[name], name + " split vs single block")
    ctx.set_math_mode("strict")


ROUT_CASES = [
    # (rout_case, l1_factor, inflow, celerity, hourly)
    (1, 1, None, None, True),
    (1, 2, None, None, False),
    (1, -2, None, None, False),
    (1, 1, (True, False), None, False),
    (2, 1, None, 0.9, False),    # TSrout < model step: sub-stepping
    (2, 2, (False,), 0.05, False),  # TSrout > model step: accumulation + back-fill
]


@pytest.mark.parametrize("rout_case,l1_factor,inflow,celerity,hourly", ROUT_CASES)
def test_routing_per_step_seam_bit_exact(ctx, rout_case, l1_factor, inflow, celerity, hourly):
    """B3: fed with the oracle's own L1 runoff, one mrm_cuda_route per routing call must be
    bit-identical to the serial netPerm sweep (all node states and the gauge series)."""
    prob = synth.make_problem(nx=24, ny=14, n_days=3, hourly=hourly, rout_case=rout_case,
                              l1_factor=l1_factor, inflow=inflow, celerity=celerity or 1.5)
    if rout_case != 1:
        orc_run.case23_params(prob["net"])
    net = prob["net"]
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob, history=True)
    ctx.set_math_mode("strict")
    dom = fresh(ctx, prob)
    idx = interface.time_indices(prob["time"], 1, prob["nTstepForcingDay"], 1, nT)
    # replay the reference's routing schedule (mo_mhm_interface_run.f90:460-514) on the host
    factor = 1.0 if rout_case == 1 else net["TSrout"] / 3600.0
    run_to_rout = np.zeros(prob["nCells"])
    inflow_acc = np.zeros(max(1, net["nInflowTotal"]))
    nI = net["nInflowTotal"]
    calls = 0
    for tt in range(1, nT + 1):
        o.run(tt, tt)
        runoff = o.F["L1_total_runoff"]
        day = (tt + 23) // 24
        qin = prob["inflowQ"][:, day - 1] if nI else np.zeros(1)
        do_route, fin, ts_rout = False, factor, 1
        if rout_case == 1 or factor < 1.0:
            run_to_rout = runoff.copy()
            inflow_acc[:] = qin if nI else 0.0
            do_route = True
        else:
            run_to_rout = run_to_rout + runoff
            inflow_acc = inflow_acc + (qin if nI else 0.0)
            if tt == nT and tt % round(fin) != 0:
                fin = float(tt % round(fin))
            if tt % round(fin) == 0 or tt == nT:
                inflow_acc = inflow_acc / fin
                ts_rout = int(round(fin))
                do_route = True
        if do_route:
            dom.route(tt, int(idx[tt - 1].yId), ts_rout, fin, RunToRout=run_to_rout,
                      InflowDischarge=inflow_acc if nI else None)
            calls += 1
            for name in ("L11_qOUT", "L11_qTIN", "L11_qTR", "L11_qMod", "L11_C1", "L11_C2"):
                if name in ("L11_C1", "L11_C2") and net["nNodes"] - net["nOutlets"] < net["nNodes"]:
                    nl = net["nNodes"] - net["nOutlets"]
                    parity.assert_bit_exact(dom.get_routing_state(name)[:nl], o.R[name][:nl],
                                            "%s @tt=%d" % (name, tt))
                else:
                    parity.assert_bit_exact(dom.get_routing_state(name), o.R[name], "%s @tt=%d" % (name, tt))
            run_to_rout = np.zeros(prob["nCells"])
            inflow_acc = np.zeros(max(1, nI))
    assert calls >= 3
    assert net["nOutlets"] >= 2, "multi-outlet quirk (only the last sink adds its runoff) untested"


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("rout_case,l1_factor,inflow,celerity,hourly", ROUT_CASES)
def test_full_run_gauge_discharge(ctx, mode, rout_case, l1_factor, inflow, celerity, hourly):
    """cells + routing in time blocks: gauge discharge over the full run <= 1e-8 relative,
    routing states <= 1e-9; uneven block split == single block (bit-exact)."""
    prob = synth.make_problem(nx=24, ny=14, n_days=6, hourly=hourly, rout_case=rout_case,
                              l1_factor=l1_factor, inflow=inflow, celerity=celerity or 1.5)
    if rout_case != 1:
        orc_run.case23_params(prob["net"])
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    ctx.set_math_mode(mode)
    dom = fresh(ctx, prob)
    dom.run_steps(1, nT)
    q = dom.get_runoff()
    worst = parity.assert_close(q, o.mRM_runoff, "mRM_runoff (%s)" % mode, rtol=parity.RTOL_Q)
    states = {}
    for name in ("L11_qOUT", "L11_qTIN", "L11_qTR", "L11_qMod"):
        states[name] = dom.get_routing_state(name)
        parity.assert_close(states[name], o.R[name], name, rtol=parity.RTOL_Q)
    print("%s: gauge discharge max relative difference %.3e" % (mode, worst))
    assert np.abs(o.mRM_runoff).max() > 0
    dom = fresh(ctx, prob)
    dom.run_steps(1, 5)
    dom.run_steps(6, 61)
    dom.run_steps(67, nT - 66)
    parity.assert_bit_exact(dom.get_runoff(), q, "gauge series split vs single block")
    for name, v in states.items():
        parity.assert_bit_exact(dom.get_routing_state(name), v, name + " split vs single block")
    ctx.set_math_mode("strict")


def test_ensemble_members_equal_single_runs(ctx):
    """nMembers parameter sets side by side == the same sets run one by one (bit-exact)."""
    prob = synth.make_problem(nx=20, ny=12, n_days=5, hourly=True)
    nT = prob["time"]["nTimeSteps"]
    rng = np.random.default_rng(11)
    members = []
    for m in range(3):
        P = synth.make_params(rng, prob["nCells"], prob["nH"], prob["nLAI"], prob["nLC"])
        P["rout_param"] = np.array(synth.ROUT1_PARAM) * rng.uniform(0.9, 1.1, 5)
        members.append(P)
    ctx.set_math_mode("strict")
    single = []
    for m in range(3):
        p1 = dict(prob)
        dom = fresh(ctx, p1, member_params=[members[m]])
        dom.run_steps(1, nT)
        single.append((dom.get_runoff(), dom.get_state("L1_soilMoist"), dom.get_routing_state("L11_qTR")))
    dom = fresh(ctx, prob, nMembers=3, member_params=members)
    dom.run_steps(1, nT)
    for m in range(3):
        parity.assert_bit_exact(dom.get_runoff(member=m), single[m][0], "gauge series member %d" % m)
        parity.assert_bit_exact(dom.get_state("L1_soilMoist", member=m), single[m][1], "soilMoist m%d" % m)
        parity.assert_bit_exact(dom.get_routing_state("L11_qTR", member=m), single[m][2], "qTR m%d" % m)
    # and against the oracle
    o = orc_run.OracleRun(prob, params={k: v for k, v in members[1].items()})
    o.run(1, nT)
    parity.assert_close(dom.get_runoff(member=1), o.mRM_runoff, "member 1 vs oracle", rtol=parity.RTOL_Q)


def test_default_state_init_and_errors(ctx):
    prob = synth.make_problem(nx=8, ny=6, n_days=2, nH=3, routing=False)
    dom = fresh(ctx, prob)
    dom.set_state("L1_snowPack", np.full(prob["nCells"], 99.0))
    dom.states_default_init(prob["horizon_depth"])
    for name, ref in prob["states0"].items():
        parity.assert_bit_exact(dom.get_state(name), ref, name)
    with pytest.raises(interface._lib.MhmCudaError, match="outside 1"):
        dom.run_steps(1, prob["time"]["nTimeSteps"] + 1)
    with pytest.raises(interface._lib.MhmCudaError, match="holds steps"):
        dom.set_meteo("pre", prob["forcing"]["pre"][:10])
        dom.run_steps(1, 24)


def test_fused_node_runoff_equals_separate_runoff_accumulation(ctx):
    """L11 == L1: the cell kernel writes the routing's node runoff itself (fused L11_runoff_acc).
    Strict mode: bit-identical to the unfused path (runoff history + qout kernel); asking for the
    runoff history switches the fusion off and returns the same series."""
    prob = synth.make_problem(nx=40, ny=30, n_days=6, hourly=True)
    nT = prob["time"]["nTimeSteps"]
    ctx.set_math_mode("strict")
    dom = fresh(ctx, prob)
    dom.run_steps(1, nT)
    q_fused = dom.get_runoff()
    with pytest.raises(RuntimeError):
        dom.get_runoff_history(nT)
    dom = fresh(ctx, prob)
    dom.keep_runoff_history(True)
    dom.run_steps(1, nT)
    parity.assert_bit_exact(dom.get_runoff(), q_fused, "gauge series fused vs unfused")
    assert dom.get_runoff_history(nT).shape == (nT, prob["nCells"])
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    parity.assert_close(q_fused, o.mRM_runoff, "gauge discharge", rtol=parity.RTOL_Q)
    ctx.set_math_mode("fast")
    dom = fresh(ctx, prob)
    dom.run_steps(1, nT)
    parity.assert_close(dom.get_runoff(), o.mRM_runoff, "gauge discharge (fast, fused)", rtol=parity.RTOL_Q)
    ctx.set_math_mode("strict")


@pytest.mark.parametrize("soil_case,pet_case,nH", [(1, -1, 2), (2, 0, 3), (4, -1, 1), (3, 0, 2)])
def test_uniform_calendar_launches_equal_general_launches(ctx, monkeypatch, soil_case, pet_case, nH):
    """Hourly forcing, fast mode: launches whose steps share yId / iLAI / month run the
    software-pipelined kernel variant without per-step calendar work and end where the calendar
    turns.  40 days from 1 January cross a month; states, last-step fluxes and the gauge series
    are bit-identical to the general kernels (MHM_CUDA_NO_UNIFORM_CALENDAR) and agree with the
    oracle -- Feddes and Jarvis, parameters in shared memory (nH <= 2) and in registers."""
    prob = synth.make_problem(nx=24, ny=14, n_days=40, hourly=True, soil_case=soil_case, pet_case=pet_case, nH=nH)
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    ctx.set_math_mode("fast")
    res = {}
    for key in ("uniform", "general"):
        if key == "general":
            monkeypatch.setenv("MHM_CUDA_NO_UNIFORM_CALENDAR", "1")
        dom = fresh(ctx, prob)
        dom.run_steps(1, 100)          # uneven calls: the node-runoff tiles are entered mid-way
        dom.run_steps(101, nT - 100)
        res[key] = {name: dom.get_variable(name) for name in STATES + FLUXES}
        res[key]["q"] = dom.get_runoff()
    monkeypatch.delenv("MHM_CUDA_NO_UNIFORM_CALENDAR")
    for name in res["uniform"]:
        parity.assert_bit_exact(res["uniform"][name], res["general"][name], name + " uniform vs general launches")
    parity.assert_close(res["uniform"]["q"], o.mRM_runoff, "gauge discharge", rtol=parity.RTOL_Q)
    for name in STATES:
        parity.assert_close(res["uniform"][name], o.S[name], name)
    ctx.set_math_mode("strict")


def test_lean_routing_kernel_equals_general_kernel(ctx, monkeypatch):
    """Levels without ghost sources / zeroed outflows run route_chain_lean_kernel (running
    offsets into the tiled histories); MHM_CUDA_NO_LEAN_ROUTING forces the general kernel.
    Gauge series and the routing state must be bit-identical, and equal the serial sweep."""
    prob = synth.make_problem(nx=70, ny=45, n_days=3, hourly=True, n_gauges=5)
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob)
    o.run(1, nT)
    ctx.set_math_mode("strict")
    res = {}
    for key in ("lean", "general"):
        if key == "general":
            monkeypatch.setenv("MHM_CUDA_NO_LEAN_ROUTING", "1")
        dom = fresh(ctx, prob)
        dom.run_steps(1, 29)       # blocks that start and end inside history tiles
        dom.run_steps(30, nT - 29)
        res[key] = {k: dom.get_routing_state(k) for k in ("L11_qTIN", "L11_qTR", "L11_qMod", "L11_qOUT")}
        res[key]["q"] = dom.get_runoff()
    monkeypatch.delenv("MHM_CUDA_NO_LEAN_ROUTING")
    for k in res["lean"]:
        parity.assert_bit_exact(res["lean"][k], res["general"][k], k + " lean vs general routing kernel")
    for k in ("L11_qTIN", "L11_qTR", "L11_qMod", "L11_qOUT"):   # the cells' runoff differs in the last bits
        parity.assert_close(res["lean"][k], o.R[k], k + " vs serial sweep", rtol=parity.RTOL_Q)
    parity.assert_close(res["lean"]["q"], o.mRM_runoff, "gauge discharge", rtol=parity.RTOL_Q)
