"""Gridded outputs accumulated on the device: mhm_cuda_set_outputs +
run_steps against the oracle's restatement of mHM_updateDataset / writeVariableTimestep, for
every window selector, every variable, split run_steps calls, ensembles -- and against the
reference's own *_mHM_Fluxes_States.nc files (tests/golden)."""
import numpy as np
import pytest

import golden_case
import orc_run
import parity
from mhm_b200 import driver, interface, synth

pytestmark = pytest.mark.gpu

ALL = np.ones(21, dtype=np.int32)
ALL[17] = 0  # neutrons: out of scope


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


def compare(dom, o, first_window, tol_exact):
    tts = dom.output_windows()
    ref = o.out_windows()[first_window:first_window + len(tts)]
    assert tts == [tt for tt, _ in ref], (tts, [tt for tt, _ in ref])
    worst = 0.0
    for w, (tt, fields) in enumerate(ref):
        for (var, hor), want in fields.items():
            got = dom.get_output(w, var, hor + 1 if hor >= 0 else 0)
            if tol_exact:
                parity.assert_close(got, want, "window %d var %d h %d" % (w, var, hor), rtol=1e-12, atol=1e-13)
            else:
                parity.assert_close(got, want, "window %d var %d h %d" % (w, var, hor))
            worst = max(worst, parity.rel_err(got, want).max())
    return len(tts), worst


@pytest.mark.parametrize("ts,hourly,soil_case,nH,warming", [(-1, True, 1, 2, 1), (-2, False, 2, 3, 0),
                                                            (-3, False, 1, 2, 3), (0, True, 4, 1, 0),
                                                            (7, True, 1, 2, 2), (100, False, 3, 2, 0)])
def test_output_windows_equal_oracle(ctx, ts, hourly, soil_case, nH, warming):
    """all 20 supported variables, every window selector, runs that straddle a year change and
    a land-cover scene change, with a warming period; the run is issued in three calls"""
    n_days = 9 if hourly else 40
    prob = synth.make_problem(nx=14, ny=9, n_days=n_days, nH=nH, hourly=hourly, soil_case=soil_case,
                              pet_case=-1 if hourly else 0, routing=False,
                              start=(1990, 12, 28) if hourly else (1990, 12, 5))
    prob["time"]["warming_days"] = warming
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob, outputs=(ALL, ts), max_windows=400)
    o.run(1, nT)
    assert o.d.out_nwin >= 1
    ctx.set_math_mode("strict")
    dom = fresh(ctx, prob)
    dom.set_outputs(ALL, ts)
    done, cuts = 0, [1, 50, 131, nT + 1]
    for a, b in zip(cuts[:-1], cuts[1:]):
        dom.run_steps(a, b - a)
        n, worst = compare(dom, o, done, tol_exact=False)
        done += n
    assert done == o.d.out_nwin
    print("ts=%d: %d windows, max rel diff %.2e" % (ts, done, worst))


def test_output_members_and_fast_mode(ctx):
    prob = synth.make_problem(nx=12, ny=8, n_days=5, hourly=True, routing=True)
    nT = prob["time"]["nTimeSteps"]
    flags = np.zeros(21, dtype=np.int32)
    flags[[2, 9, 10, 15]] = 1  # SWC, aET, Q, recharge
    rng = np.random.default_rng(3)
    members = [prob["params"], {k: (v * rng.uniform(0.95, 1.05) if k in ("L1_kPerco", "L1_alpha") else v)
                                for k, v in prob["params"].items()}]
    refs = []
    for P in members:
        o = orc_run.OracleRun(prob, params=P, outputs=(flags, -1))
        o.run(1, nT)
        refs.append(o)
    for mode in ("strict", "fast"):
        ctx.set_math_mode(mode)
        dom = fresh(ctx, prob, nMembers=2, member_params=members)
        dom.set_outputs(flags, -1)
        dom.run_steps(1, nT)
        tts = dom.output_windows()
        assert tts == [tt for tt, _ in refs[0].out_windows()]
        for m, o in enumerate(refs):
            for w, (tt, fields) in enumerate(o.out_windows()):
                for (var, hor), want in fields.items():
                    got = dom.get_output(w, var, hor + 1 if hor >= 0 else 0, member=m)
                    parity.assert_close(got, want, "%s member %d window %d var %d" % (mode, m, w, var))
        parity.assert_close(dom.get_runoff(member=1), refs[1].mRM_runoff, "gauge discharge", rtol=parity.RTOL_Q)
    ctx.set_math_mode("strict")


@pytest.mark.parametrize("case", ["case_00", "case_09", "case_04_b2"])
def test_outputs_reproduce_reference_fluxes_states_file(ctx, case):
    """the reference's own gridded output (means / sums over the whole evaluation period)"""
    prob, ref = golden_case.load(case)
    if prob["rout_case"] == 2:
        orc_run.case23_params(prob["net"])
    outs = ref["outputs"]
    ctx.set_math_mode("strict")
    dom = fresh(ctx, prob)
    dom.set_outputs(outs["flags"], outs["timestep"])
    dom.run_steps(1, prob["time"]["nTimeSteps"])
    assert dom.output_windows() == [prob["time"]["nTimeSteps"]]
    for (var, hor), want in outs["fields"].items():
        got = dom.get_output(0, var, hor + 1 if hor >= 0 else 0)
        worst = parity.assert_close(got, want[0], "%s variable %d horizon %d" % (case, var, hor))
        print("%s var %d h %d: max rel diff %.2e" % (case, var, hor, worst))


@pytest.mark.parametrize("nH,soil_case", [(2, 1), (1, 4), (2, 2)])
def test_default_output_set_in_registers_equals_shared_memory_path(ctx, monkeypatch, nH, soil_case):
    """The reference's default mhm_outputs.nml (variables 1-16, 19-21, monthly) on hourly forcing in
    fast mode runs the kernel that keeps the open window in registers (OUT = 2); the general kernel
    (window in shared memory, run-time selection; MHM_CUDA_NO_OUTPUT_REGISTERS) must give the same
    windows bit for bit, both equal the oracle, with a warming period, a year change inside the run and
    the run issued in uneven calls."""
    flags = np.zeros(21, dtype=np.int32)
    flags[:16] = 1
    flags[18:21] = 1
    prob = synth.make_problem(nx=30, ny=17, n_days=40, nH=nH, hourly=True, soil_case=soil_case, pet_case=-1,
                              routing=False, start=(1990, 12, 10))
    prob["time"]["warming_days"] = 2
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob, outputs=(flags, -2))
    o.run(1, nT)
    ctx.set_math_mode("fast")
    res = {}
    for key in ("registers", "shared"):
        if key == "shared":
            monkeypatch.setenv("MHM_CUDA_NO_OUTPUT_REGISTERS", "1")
        dom = fresh(ctx, prob, nMembers=2, member_params=[prob["params"]] * 2)
        dom.set_outputs(flags, -2)
        got, first = {}, 0
        for a, b in ((1, 300), (301, 331), (632, nT - 631)):
            dom.run_steps(a, b)
            nwin, _ = compare(dom, o, first, tol_exact=False)
            for w, tt in enumerate(dom.output_windows()):
                for sl, (var, hor) in enumerate(o.out_windows()[first + w][1]):
                    got[(tt, var, hor)] = dom.get_output(w, var, hor + 1 if hor >= 0 else 0, member=1)
            first += nwin
        assert first == len(o.out_windows()) >= 2
        res[key] = got
    monkeypatch.delenv("MHM_CUDA_NO_OUTPUT_REGISTERS")
    assert res["registers"].keys() == res["shared"].keys()
    for k in res["registers"]:
        parity.assert_bit_exact(res["registers"][k], res["shared"][k], "window ending %d, variable %d, horizon %d" % k)
    ctx.set_math_mode("strict")


@pytest.mark.parametrize("nH,soil_case,routing", [(2, 1, True), (1, 4, True), (2, 2, False)])
def test_default_outputs_on_pipelined_uniform_launches_equal_general_launches(ctx, monkeypatch, nH, soil_case,
                                                                               routing):
    """Launches that accumulate the default output set are uniform-calendar launches too (fast mode):
    they run the software-pipelined kernel with the output sums split by stage (accumulate_default_a
    / _b), with the fused node-runoff store when the domain is routed and, with
    mhm_cuda_keep_runoff_history, with the total-runoff history.  MHM_CUDA_NO_UNIFORM_OUTPUTS keeps
    them on the general per-step path: the windows, the final states and the gauge series must be
    the same bit for bit, and the windows equal the oracle's."""
    flags = np.zeros(21, dtype=np.int32)
    flags[:16] = 1
    flags[18:21] = 1
    prob = synth.make_problem(nx=30, ny=17, n_days=40, nH=nH, hourly=True, soil_case=soil_case, pet_case=-1,
                              routing=routing, start=(1990, 12, 10))
    prob["time"]["warming_days"] = 2
    nT = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob, outputs=(flags, -2))
    o.run(1, nT)
    ctx.set_math_mode("fast")
    res = {}
    for key in ("pipelined", "general"):
        if key == "general":
            monkeypatch.setenv("MHM_CUDA_NO_UNIFORM_OUTPUTS", "1")
        dom = fresh(ctx, prob, nMembers=2, member_params=[prob["params"]] * 2)
        if not routing:
            dom.keep_runoff_history(True)
        dom.set_outputs(flags, -2)
        got, first = {}, 0
        for a, b in ((1, 300), (301, 331), (632, nT - 631)):
            dom.run_steps(a, b)
            nwin, _ = compare(dom, o, first, tol_exact=False)
            for w, tt in enumerate(dom.output_windows()):
                for sl, (var, hor) in enumerate(o.out_windows()[first + w][1]):
                    got[(tt, var, hor)] = dom.get_output(w, var, hor + 1 if hor >= 0 else 0, member=1)
            first += nwin
        assert first == len(o.out_windows()) >= 2
        for name in ("L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"):
            got[("state", name, 0)] = dom.get_state(name, member=1)
        if routing:
            got[("gauge", 0, 0)] = dom.get_runoff(1, nT, member=1)
        res[key] = got
    monkeypatch.delenv("MHM_CUDA_NO_UNIFORM_OUTPUTS")
    assert res["pipelined"].keys() == res["general"].keys()
    for k in res["pipelined"]:
        parity.assert_bit_exact(res["pipelined"][k], res["general"][k], "%s %s %s" % k)
    ctx.set_math_mode("strict")
