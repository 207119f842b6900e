"""The whole path from the reference's RAW inputs on this stack, against the reference's run:
ASCII morphology + look-up tables + gamma (mhm_parameter.nml) + daily forcing
 -> init_lowres_level (L1, L11), L11_L1_mapping, mrm_net_init (river network, flood plains)
 -> mpr_cuda_eval (all L1 effective parameters on the device)
 -> cascade + routing + gridded outputs on the device (run_steps)
 -> daily gauge discharge, final states and the Fluxes_States output of check/case_*.
Nothing the reference computed is fed in (the restart files are only compared against)."""
import datetime

import numpy as np
import pytest

import golden_case
import parity
from mhm_b200 import interface, netinit, synth, synth_mpr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,res1,res11", [("case_00", 24000.0, 24000.0), ("case_04_b2", 12000.0, 24000.0),
                                             ("case_04_b5", 12000.0, 12000.0)])
def test_raw_inputs_to_discharge(case, res1, res11):
    z0 = np.load(golden_case.HERE + "/golden/test_domain_l0.npz")
    zc = np.load(golden_case.HERE + "/golden/%s.npz" % case)
    prob_ref, ref = golden_case.load(case)      # forcing + the reference's results
    mprob, _ = golden_case.load_mpr(case, synth_mpr.init_lowres_level)
    cs0 = float(z0["cellsize0"])
    n0 = int(z0["mask0"].sum())
    g1 = mprob["grid"]
    g11 = synth_mpr.init_lowres_level(z0["mask0"], cs0, res11, np.full(n0, cs0 * cs0))
    net0 = netinit.net_init(z0["mask0"], z0["fDir0"], z0["fAcc0"], z0["elev0"], cs0, g11, z0["gaugeLoc0"], [398],
                            LCover0=z0["LCover0"])
    l1_l11, _ = netinit.l1_l11_mapping(g1, res1, g11, res11)
    nn, nl = g11["nCells1"], net0["nLinks"]
    net = {"nNodes": nn, "nOutlets": nn - nl, "map_flag": 1, "fromN": net0["fromN"], "toN": net0["toN"],
           "netPerm": net0["netPerm"], "L1_L11_Id": l1_l11, "L11_L1_Id": np.ones(nn, dtype=np.int32),
           "L1_areaCell": g1["cellArea1"] * 1e-6, "L11_areaCell": g11["cellArea1"] * 1e-6,
           "gaugeNodeList": net0["gaugeNodeList"], "gaugeIndexList": np.array([1], dtype=np.int32),
           "nGaugesTotal": 1, "processCase": 1}
    n1 = g1["nCells1"]
    with interface.Context() as ctx:
        ctx.set_math_mode("strict")
        dom = ctx.register_domain(1, n1, mprob["nH"], mprob["nLAI"], mprob["nLC"], mprob["processMatrix"])
        synth_mpr.set_mpr_inputs(dom, mprob)
        dom.set_meteo_config(prob_ref["pet_case"], 1, False, False, synth.FNIGHT_PREC, synth.FNIGHT_PET,
                             synth.FNIGHT_TEMP, synth.EVAP_COEFF)
        dom.set_time(prob_ref["time"])
        synth_mpr.mpr_eval(dom, mprob["param"])                      # gamma -> L1 parameters, on the device
        dom.states_default_init(np.array([200.0, 1000.0]))
        for var in ("pre", "temp", "pet"):
            dom.set_meteo(var, prob_ref["forcing"][var])
        dom.set_network(net)
        dom.set_reg_rout(zc["rout_param"], net0["length"][: nn - 1], net0["slope"][: nn - 1], net0["nLinkFracFPimp"])
        outs = ref["outputs"]
        dom.set_outputs(outs["flags"], outs["timestep"])
        dom.run_steps(1, prob_ref["time"]["nTimeSteps"])
        q = golden_case.daily_mean(dom.get_runoff(), ref["warming_days"])
        wq = parity.assert_close(q, ref["Qsim"], case + " daily discharge", rtol=parity.RTOL_Q)
        worst = 0.0
        for name in ("L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"):
            worst = max(worst, parity.assert_close(dom.get_variable(name), ref["final"][name], case + " " + name))
        for ours, theirs in (("L11_qTIN", "L11_qTIN"), ("L11_qTR", "L11_qTR"), ("L11_C1", "L11_C1"), ("L11_C2", "L11_C2")):
            parity.assert_close(dom.get_routing_state(ours), ref["final"][theirs], case + " " + theirs, rtol=parity.RTOL_Q)
        for (var, hor), want in outs["fields"].items():
            parity.assert_close(dom.get_output(0, var, hor + 1 if hor >= 0 else 0), want[0], "%s output %d" % (case, var))
        print("%s from raw inputs: daily discharge max rel diff %.2e, states %.2e" % (case, wq, worst))


def test_raw_inputs_to_discharge_celerity_routing():
    """check/case_13 (routing processCase 3): river network, link-length floor, L11_calc_celerity
    and mrm_update_param from the raw grids, MPR + cascade + adaptive-step routing on the device,
    against the reference's daily discharge and final states"""
    case = "case_13"
    z0 = np.load(golden_case.HERE + "/golden/test_domain_l0.npz")
    zc = np.load(golden_case.HERE + "/golden/%s.npz" % case)
    prob_ref, ref = golden_case.load(case)
    mprob, _ = golden_case.load_mpr(case, synth_mpr.init_lowres_level)
    cs0 = float(z0["cellsize0"])
    n0 = int(z0["mask0"].sum())
    g1 = mprob["grid"]
    g11 = synth_mpr.init_lowres_level(z0["mask0"], cs0, 24000.0, np.full(n0, cs0 * cs0))
    net0 = netinit.net_init(z0["mask0"], z0["fDir0"], z0["fAcc0"], z0["elev0"], cs0, g11, z0["gaugeLoc0"], [398],
                            LCover0=z0["LCover0"], routingCase=3)
    l1_l11, _ = netinit.l1_l11_mapping(g1, 24000.0, g11, 24000.0)
    nn, nl = g11["nCells1"], net0["nLinks"]
    cel11, _ = netinit.calc_celerity(z0["mask0"], z0["fDir0"], z0["slope0"], net0, float(zc["slope_factor"][0]))
    c1, c2, ts = netinit.update_param(net0["length"], cel11, nn - nl)
    net = {"nNodes": nn, "nOutlets": nn - nl, "map_flag": 1, "fromN": net0["fromN"], "toN": net0["toN"],
           "netPerm": net0["netPerm"], "L1_L11_Id": l1_l11, "L11_L1_Id": np.ones(nn, dtype=np.int32),
           "L1_areaCell": g1["cellArea1"] * 1e-6, "L11_areaCell": g11["cellArea1"] * 1e-6,
           "gaugeNodeList": net0["gaugeNodeList"], "gaugeIndexList": np.array([1], dtype=np.int32),
           "nGaugesTotal": 1, "processCase": 3}
    for mode in ("strict", "fast"):
        with interface.Context() as ctx:
            ctx.set_math_mode(mode)
            dom = ctx.register_domain(1, g1["nCells1"], mprob["nH"], mprob["nLAI"], mprob["nLC"], mprob["processMatrix"])
            synth_mpr.set_mpr_inputs(dom, mprob)
            dom.set_meteo_config(prob_ref["pet_case"], 1, False, False, synth.FNIGHT_PREC, synth.FNIGHT_PET,
                                 synth.FNIGHT_TEMP, synth.EVAP_COEFF)
            dom.set_time(prob_ref["time"])
            synth_mpr.mpr_eval(dom, mprob["param"])
            dom.states_default_init(np.array([200.0, 1000.0]))
            for var in ("pre", "temp", "pet"):
                dom.set_meteo(var, prob_ref["forcing"][var])
            dom.set_network(net)
            dom.set_c1c2(c1, c2, ts)
            dom.run_steps(1, prob_ref["time"]["nTimeSteps"])
            q = golden_case.daily_mean(dom.get_runoff(), ref["warming_days"])
            wq = parity.assert_close(q, ref["Qsim"], case + " daily discharge", rtol=parity.RTOL_Q)
            worst = 0.0
            for name in ("L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"):
                worst = max(worst, parity.assert_close(dom.get_variable(name), ref["final"][name], case + " " + name))
            print("%s[%s] from raw inputs (TSrout %g s): daily discharge max rel diff %.2e, states %.2e" % (
                case, mode, ts, wq, worst))


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_penman_monteith_calibration_run(mode):
    """check/case_03 (PET processCase 3, final run of a DDS calibration): MPR incl. aerodynamic
    and bulk surface resistance, Penman-Monteith prologue, cascade and routing on the device
    against the reference's saved discharge"""
    from test_golden_reference import _case03
    from mhm_b200 import driver

    def device_mpr(mprob):
        with interface.Context() as c:
            c.set_math_mode("strict")
            dom = c.register_domain(1, mprob["nL1"], mprob["nH"], mprob["nLAI"], mprob["nLC"], mprob["processMatrix"])
            synth_mpr.set_mpr_inputs(dom, mprob)
            synth_mpr.mpr_eval(dom, mprob["param"])
            out = {}
            for name in synth_mpr.outputs_for(1, 3):
                d2, d3 = synth_mpr.MPR_OUTPUTS[name](mprob["nH"], mprob["nLAI"], mprob["nLC"])
                out[name] = dom.get_param(name, d2, d3)
            return out

    prob, ref = _case03(synth_mpr.init_lowres_level, device_mpr)
    with interface.Context() as ctx:
        ctx.set_math_mode(mode)
        dom = driver.setup_domain(ctx, 1, prob)
        dom.run_steps(1, prob["time"]["nTimeSteps"])
        q = golden_case.daily_mean(dom.get_runoff(), ref["warming_days"])
        worst = parity.assert_close(q, ref["Qsim"], "case_03 daily discharge (%s)" % mode, rtol=parity.RTOL_Q)
        print("case_03[%s]: daily discharge max rel diff %.2e" % (mode, worst))
