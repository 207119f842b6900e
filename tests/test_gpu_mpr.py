"""GPU parity of the MPR path (upscaling operators, gamma -> L1 parameters) against the oracle."""
import ctypes as C

import numpy as np
import pytest

import orc_mpr
import orc_run
import parity
from mhm_b200 import _lib, driver, interface, synth, synth_mpr
from mhm_b200.interface import _pd, _pi, check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = interface.Context()
    yield c
    c.finalize()


def make_grid(ctx, prob):
    g = prob["grid"]
    h = C.c_void_p()
    check(ctx.L.mpr_cuda_grid_create(ctx.h, prob["nrows0"], prob["ncols0"], _pi(prob["mask0"]), prob["nL1"],
                                     _pi(g["upper_bound"]), _pi(g["lower_bound"]), _pi(g["left_bound"]),
                                     _pi(g["right_bound"]), _pi(g["n_subcells"]), C.byref(h)))
    return h


@pytest.mark.parametrize("factor", [1, 4, 7])
def test_upscaling_operators(ctx, factor):
    prob = synth_mpr.make_mpr_problem(nx0=90, ny0=70, factor=factor)
    rng = np.random.default_rng(3)
    x = np.ascontiguousarray(rng.uniform(0.5, 3.0, prob["nL0"]))
    cls = rng.integers(1, 4, prob["nL0"]).astype(np.int32)
    grid = make_grid(ctx, prob)
    out = np.zeros(prob["nL1"])
    calls = {"arith": ctx.L.mpr_cuda_upscale_arithmetic_mean, "harm": ctx.L.mpr_cuda_upscale_harmonic_mean,
             "geom": ctx.L.mpr_cuda_upscale_geometric_mean}
    for mode in ("strict", "fast"):
        ctx.set_math_mode(mode)
        for op, fn in calls.items():
            check(fn(ctx.h, grid, -9999.0, _pd(x), _pd(out)))
            ref = orc_mpr.upscale(prob, op, x)
            if mode == "strict" and op != "geom":
                parity.assert_bit_exact(out, ref, "%s (strict = serial order)" % op)
            else:
                parity.assert_close(out, ref, "%s (%s)" % (op, mode), rtol=1e-12, atol=0)
        check(ctx.L.mpr_cuda_l0_fractional_cover(ctx.h, grid, _pi(cls), 2, _pd(out)))
        parity.assert_bit_exact(out, orc_mpr.upscale(prob, "frac", cls, class_id=2), "fractional cover")
    ctx.set_math_mode("strict")
    check(ctx.L.mpr_cuda_grid_destroy(ctx.h, grid))


def register(ctx, prob):
    for k in list(ctx.domains):
        check(ctx.L.mhm_cuda_unregister_domain(ctx.h, k))
        del ctx.domains[k]
    dom = ctx.register_domain(1, prob["nL1"], prob["nH"], prob["nLAI"], prob["nLC"], prob["processMatrix"])
    synth_mpr.set_mpr_inputs(dom, prob)
    return dom


MPR_CASES = [(1, -1, 2), (2, 0, 3), (3, 1, 2), (4, 2, 3), (1, 3, 2), (2, -1, 4)]


@pytest.mark.parametrize("soil_case,pet_case,nH", MPR_CASES)
def test_mpr_eval_equals_oracle(ctx, soil_case, pet_case, nH):
    """all L1 effective parameters for every process variant; strict mode.  Fields that involve
    device exp/log/pow (petLAIcorFactor, aeroResist, FC-dependent roots) within 1e-12, all others
    are pure adds/mults/divisions in the reference's order and must be bit-identical."""
    prob = synth_mpr.make_mpr_problem(nx0=80, ny0=50, factor=5, nH=nH, soil_case=soil_case, pet_case=pet_case)
    ref = orc_mpr.run_mpr(prob)
    ctx.set_math_mode("strict")
    dom = register(ctx, prob)
    synth_mpr.mpr_eval(dom, prob["param"])
    loose = {"L1_petLAIcorFactor", "L1_aeroResist"}
    if soil_case in (3, 4):
        loose.add("L1_fRoots")
    for name in synth_mpr.outputs_for(soil_case, pet_case):
        d2, d3 = synth_mpr.MPR_OUTPUTS[name](nH, prob["nLAI"], prob["nLC"])
        got = dom.get_param(name, d2, d3)
        if name in loose:
            parity.assert_close(got, ref[name], name, rtol=1e-12, atol=0)
        else:
            parity.assert_bit_exact(got, ref[name], name)
    # fast mode (warp-shuffle reductions): <= 1e-12 relative
    ctx.set_math_mode("fast")
    synth_mpr.mpr_eval(dom, prob["param"])
    for name in synth_mpr.outputs_for(soil_case, pet_case):
        d2, d3 = synth_mpr.MPR_OUTPUTS[name](nH, prob["nLAI"], prob["nLC"])
        parity.assert_close(dom.get_param(name, d2, d3), ref[name], name + " (fast)", rtol=1e-11, atol=0)
    ctx.set_math_mode("strict")


def test_mpr_then_cells_end_to_end(ctx):
    """gamma -> MPR -> cascade entirely on the device vs oracle MPR -> oracle cascade."""
    mp = synth_mpr.make_mpr_problem(nx0=80, ny0=50, factor=5, nH=2, soil_case=1, pet_case=-1)
    n1 = mp["nL1"]
    ref_params = orc_mpr.run_mpr(mp)
    prob = synth.make_problem(nx=8, ny=8, n_days=5, hourly=True, routing=False)
    # graft the MPR domain onto the synthetic forcing problem
    rng = np.random.default_rng(1)
    prob["nCells"] = n1
    prob["processMatrix"] = mp["processMatrix"]
    prob["params"] = dict(ref_params)  # unused fields stay zero, like variables_default_init leaves them
    nT = prob["time"]["nTimeSteps"]
    prob["forcing"] = synth.make_forcing(rng, n1, nT, True, -1)
    prob["horizon_depth"] = np.array([200.0, 400.0])
    prob["states0"] = synth.default_states(n1, 2, prob["horizon_depth"])
    o = orc_run.OracleRun(prob, history=True)
    o.run(1, nT)
    ctx.set_math_mode("strict")
    dom = register(ctx, mp)
    dom.set_meteo_config(-1, 24, True, False, synth.FNIGHT_PREC, synth.FNIGHT_PET, synth.FNIGHT_TEMP,
                         synth.EVAP_COEFF)
    dom.set_time(prob["time"])
    synth_mpr.mpr_eval(dom, mp["param"])
    dom.states_default_init(prob["horizon_depth"])
    for var in ("pre", "temp", "pet"):
        dom.set_meteo(var, prob["forcing"][var])
    dom.run_steps(1, nT)
    hist = dom.get_runoff_history(nT)
    ref = np.stack([o.hist("L1_total_runoff", tt) for tt in range(1, nT + 1)])
    worst = parity.assert_close(hist, ref, "total runoff history after device MPR")
    print("max relative difference %.2e" % worst)
    assert ref.max() > 0


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_mpr_sharded_by_l1_cells_equals_whole_domain(ctx, mode):
    """BASELINE config 4: the domain's L1 cells are dealt to shards (here three irregular sets, like
    sub-catchments); every shard gets only the L0 cells under its own L1 cells
    (synth_mpr.shard_mpr_problem) and evaluates MPR on them.  Its parameters equal the whole domain's
    on the same cells bit for bit -- an L1 cell sees only its own L0 rectangle -- and the oracle
    agrees on the shard's inputs."""
    prob = synth_mpr.make_mpr_problem(nx0=96, ny0=70, factor=6, nH=2, soil_case=1, pet_case=-1)
    ctx.set_math_mode(mode)
    dom = register(ctx, prob)
    synth_mpr.mpr_eval(dom, prob["param"])
    names = synth_mpr.outputs_for(1, -1)
    shape = {n: synth_mpr.MPR_OUTPUTS[n](2, prob["nLAI"], prob["nLC"]) for n in names}
    whole = {n: dom.get_param(n, *shape[n]) for n in names}
    rng = np.random.default_rng(5)
    owner = rng.integers(0, 3, prob["nL1"])
    owner[: prob["nL1"] // 4] = 0          # one contiguous stretch and scattered cells
    seen = 0
    for r in range(3):
        cells = np.nonzero(owner == r)[0]
        sub = synth_mpr.shard_mpr_problem(prob, cells)
        assert sub["nL0"] < prob["nL0"] and sub["nL0"] == int(np.asarray(prob["grid"]["n_subcells"])[cells].sum())
        seen += sub["nL0"]
        sd = ctx.register_domain(10 + r, sub["nL1"], sub["nH"], sub["nLAI"], sub["nLC"], sub["processMatrix"])
        synth_mpr.set_mpr_inputs(sd, sub)
        synth_mpr.mpr_eval(sd, sub["param"])
        for n in names:
            parity.assert_bit_exact(sd.get_param(n, *shape[n]), np.ascontiguousarray(whole[n][..., cells]),
                                    "shard %d: %s" % (r, n))
        if mode == "strict" and r == 1:
            ref = orc_mpr.run_mpr(sub)
            for n in names:
                if n == "L1_petLAIcorFactor":
                    parity.assert_close(sd.get_param(n, *shape[n]), ref[n], n, rtol=1e-12, atol=0)
                else:
                    parity.assert_bit_exact(sd.get_param(n, *shape[n]), ref[n], "oracle on the shard: " + n)
    assert seen == prob["nL0"]
    ctx.set_math_mode("strict")
