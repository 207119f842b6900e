"""The reference's OWN end-to-end outputs as golden vectors (tests/golden/case_*.npz, generated
by tests/golden/make_golden.py from /root/reference/check/case_*/output_save and the test
domain's forcing files).

Every fixture holds the inputs of the hot path as the Fortran run saw them (forcing, the
effective parameters its MPR wrote to the restart file, the river network) and its results
after 364..911 simulated days of hourly steps: final states and last-step fluxes, the final
routing state with C1/C2, and the daily gauge discharge in double precision.

  * CPU (`-m "not gpu"`): the oracle must reproduce the reference -- this is what pins the
    oracle (cascade, meteo day/night disaggregation, PET cases -1/0/1/2, calendar, land-cover
    scene switch, runoff accumulation with 1 and 4 L1 cells per node, Muskingum routing cases
    1 and 2, reg_rout).  Measured: final states, fluxes and routing state bit-identical to
    the Fortran run in every case; daily discharge within 1 ulp of the daily mean.
  * GPU (`-m gpu`): the CUDA path through the C ABI against the same reference outputs at the
    north-star tolerances (states/fluxes 1e-9, gauge discharge 1e-8).
"""
import numpy as np
import pytest

import golden_case
import orc_run
import parity

CASES = ["case_00", "case_02", "case_09", "case_10", "case_12", "case_13",
         "case_04_b1", "case_04_b2", "case_04_b4", "case_04_b5"]
STATES = ["L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW", "L1_soilMoist"]
FLUXES = ["L1_aETCanopy", "L1_aETSealed", "L1_baseflow", "L1_fastRunoff", "L1_melt", "L1_percol",
          "L1_preEffect", "L1_rain", "L1_runoffSeal", "L1_slowRunoff", "L1_snow", "L1_Throughfall",
          "L1_total_runoff", "L1_aETSoil", "L1_infilSoil"]
ROUT = {"L11_C1": "L11_C1", "L11_C2": "L11_C2", "L11_qOUT": "L11_qOUT", "L11_qTIN": "L11_qTIN",
        "L11_qTR": "L11_qTR", "L11_qMod": "L11_Qmod"}


def _load(case):
    prob, ref = golden_case.load(case)
    if prob["rout_case"] == 2:
        orc_run.case23_params(prob["net"])   # mrm_update_param (constant celerity) -> C1, C2, TSrout
        assert prob["net"]["TSrout"] == int(np.load(golden_case.HERE + "/golden/%s.npz" % case)["net/L11_TSrout"][0])
    if prob["rout_case"] == 3:
        # check/case_13: celerity from the L0 river slopes (L11_calc_celerity incl. FORCES mad),
        # link-length floor, mrm_update_param -- the run saved no mRM restart, its daily
        # discharge pins all of it
        orc_run.case3_params(prob["net"], np.load(golden_case.HERE + "/golden/test_domain_l0.npz"))
    return prob, ref


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_run(case):
    prob, ref = _load(case)
    o = orc_run.OracleRun(prob)
    o.run(1, prob["time"]["nTimeSteps"])
    for name in STATES:
        parity.assert_bit_exact(o.S[name], ref["final"][name], "%s %s" % (case, name))
    for name in FLUXES:
        parity.assert_bit_exact(o.F[name], ref["final"][name], "%s %s" % (case, name))
    if prob["net"] is None:
        return
    for ours, theirs in ROUT.items():
        if theirs in ref["final"]:
            parity.assert_bit_exact(o.R[ours], ref["final"][theirs], "%s %s" % (case, theirs))
    q = golden_case.daily_mean(o.mRM_runoff, ref["warming_days"])
    # the daily mean is a sum of 24 values: summation order of the Fortran `sum` intrinsic
    # (gfortran may vectorise it) is the only freedom left -> 4 ulp
    worst = parity.assert_close(q, ref["Qsim"], case + " daily discharge", rtol=1e-15, atol=0.0)
    if ref["Qsim_text"] is not None:
        assert np.abs(q - ref["Qsim_text"]).max() < 0.6e-7, "7-decimal text table"
    print("%s: daily discharge max rel diff %.2e over %d days" % (case, worst, q.shape[1]))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_reproduces_reference_run(case, mode):
    from mhm_b200 import driver, interface

    prob, ref = _load(case)
    with interface.Context() as ctx:
        ctx.set_math_mode(mode)
        dom = driver.setup_domain(ctx, 1, prob)
        dom.run_steps(1, prob["time"]["nTimeSteps"])
        worst = 0.0
        for name in STATES + FLUXES:
            worst = max(worst, parity.assert_close(dom.get_variable(name), ref["final"][name],
                                                   "%s %s (%s)" % (case, name, mode)))
        msg = "%s[%s]: states/fluxes max rel diff %.2e" % (case, mode, worst)
        if prob["net"] is not None:
            for ours, theirs in ROUT.items():
                if theirs in ref["final"]:
                    parity.assert_close(dom.get_routing_state(ours), ref["final"][theirs],
                                        "%s %s (%s)" % (case, theirs, mode), rtol=parity.RTOL_Q)
            q = golden_case.daily_mean(dom.get_runoff(), ref["warming_days"])
            wq = parity.assert_close(q, ref["Qsim"], case + " daily discharge", rtol=parity.RTOL_Q)
            msg += ", daily discharge %.2e" % wq
        print(msg)


# ---- MPR: gamma -> L1 effective parameters on the real test basin ------------------------------
MPR_CASES = ["case_00", "case_02", "case_09", "case_10", "case_12", "case_04_b2"]


@pytest.mark.parametrize("case", MPR_CASES)
def test_oracle_mpr_reproduces_reference_parameters(case):
    """L0 fields of the bundled test basin (tests/golden/test_domain_l0.npz, 46 545 L0 cells,
    1 475 soil types) + the case's gamma -> every L1 effective parameter the reference's MPR wrote
    to its restart file, bit for bit; also pins init_lowres_level (L1 mask and cell areas at
    24 km and 12 km) for the oracle and for the library's host helper."""
    import orc_mpr
    from mhm_b200 import synth_mpr

    prob, ref = golden_case.load_mpr(case, orc_mpr.init_lowres_level)
    g = prob["grid"]
    assert np.array_equal(g["mask1"] != 0, ref["mask1"])
    parity.assert_bit_exact(g["cellArea1"] * 1e-6, ref["L1_areaCell_km2"], "L1 cell area")
    lib = synth_mpr.init_lowres_level(prob["mask0"], 500.0, 500.0 * 432 // ref["mask1"].shape[0],
                                      np.full(prob["nL0"], 250000.0))
    for k in ("mask1", "cellArea1", "upper_bound", "lower_bound", "left_bound", "right_bound",
              "n_subcells", "lowres_id_on_highres"):
        assert np.array_equal(lib[k], g[k]), k
    out = orc_mpr.run_mpr(prob)
    checked = 0
    for name, want in ref.items():
        if name in out:
            parity.assert_bit_exact(out[name], want, "%s %s" % (case, name))
            checked += 1
    assert checked >= 20


@pytest.mark.gpu
@pytest.mark.parametrize("case", MPR_CASES)
def test_cuda_mpr_reproduces_reference_parameters(case):
    """mpr_cuda_eval on the device against the reference's restart parameters: strict mode is
    bit-identical except where device pow/exp/log enter (<= 1e-12); fast mode <= 1e-11."""
    import orc_mpr
    from mhm_b200 import interface, synth_mpr

    prob, ref = golden_case.load_mpr(case, synth_mpr.init_lowres_level)
    loose = {"L1_petLAIcorFactor", "L1_aeroResist"} | ({"L1_fRoots"} if prob["soil_case"] in (3, 4) else set())
    with interface.Context() as ctx:
        dom = ctx.register_domain(1, prob["nL1"], prob["nH"], prob["nLAI"], prob["nLC"], prob["processMatrix"])
        synth_mpr.set_mpr_inputs(dom, prob)
        for mode in ("strict", "fast"):
            ctx.set_math_mode(mode)
            synth_mpr.mpr_eval(dom, prob["param"])
            for name in synth_mpr.outputs_for(prob["soil_case"], prob["pet_case"]):
                d2, d3 = synth_mpr.MPR_OUTPUTS[name](prob["nH"], prob["nLAI"], prob["nLC"])
                got = dom.get_param(name, d2, d3)
                if mode == "strict" and name not in loose:
                    parity.assert_bit_exact(got, ref[name], "%s %s" % (case, name))
                else:
                    parity.assert_close(got, ref[name], "%s %s (%s)" % (case, name, mode),
                                        rtol=1e-12 if mode == "strict" else 1e-11, atol=0)


@pytest.mark.parametrize("case", ["case_00", "case_02", "case_09", "case_04_b2", "case_04_b5"])
def test_oracle_reproduces_reference_gridded_outputs(case):
    """mHM_updateDataset + writeVariableTimestep restated in the oracle against the reference's
    *_mHM_Fluxes_States.nc: soil water content per horizon (window mean), PET, aET (derived with
    the scene of the NEXT step), total runoff, recharge (window sums) -- bit-identical."""
    prob, ref = _load(case)
    outs = ref["outputs"]
    o = orc_run.OracleRun(prob, outputs=(outs["flags"], outs["timestep"]))
    o.run(1, prob["time"]["nTimeSteps"])
    W = o.out_windows()
    assert [tt for tt, _ in W] == [prob["time"]["nTimeSteps"]]
    assert set(W[0][1]) == set(outs["fields"])
    for key, want in outs["fields"].items():
        parity.assert_bit_exact(W[0][1][key], want[0], "%s output %s" % (case, key))


# ---- Penman-Monteith PET (processCase(5) = 3): the final run of a DDS calibration -------------
def _case03(init_lowres_level, run_mpr):
    """check/case_03: gamma = the optimiser's best set (FinalParam.out, 15 digits); the L1
    parameters come from MPR on the raw test-basin inputs, forcing incl. net radiation, vapour
    pressure and wind speed from the reference's files"""
    mprob, _ = golden_case.load_mpr("case_03", init_lowres_level)
    params = run_mpr(mprob)
    z = np.load(golden_case.HERE + "/golden/case_03.npz")
    # golden_case.load needs param/* entries: build the problem by hand from the MPR output
    import datetime
    from mhm_b200 import synth
    n = mprob["nL1"]
    ordinal0, n_days, warming = [int(x) for x in z["time"]]
    d0 = datetime.date.fromordinal(ordinal0)
    lc = z["lc_years"]
    prob = {"nH": 2, "nLAI": 12, "nLC": 2, "timestep_h": 1, "hourly": False, "soil_case": 1, "pet_case": 3,
            "rout_case": 1, "read_weights": False, "nCells": n, "nTstepForcingDay": 1,
            "processMatrix": z["processMatrix"],
            "time": {"jul_start": synth.JUL_1990_01_01 + (d0 - datetime.date(1990, 1, 1)).days,
                     "nTimeSteps": n_days * 24, "warming_days": warming, "timeStep_LAI_input": 0,
                     "lc_year_start": int(lc[0]), "LCyearId": np.asarray(lc[1:], dtype=np.int32)}}
    full = synth.make_params(np.random.default_rng(0), n, 2, 12, 2, 3)
    P = {k: np.ascontiguousarray(v) for k, v in params.items()}
    P["latitude"] = z["L1_lat"][None, None, :]
    for k, v in full.items():
        P.setdefault(k, v)
    prob["params"] = P
    prob["forcing"] = {k[8:]: np.ascontiguousarray(z[k]) for k in z.files if k.startswith("forcing/")}
    prob["horizon_depth"] = np.ascontiguousarray(z["horizon_bnds"][:, 1])
    prob["states0"] = synth.default_states(n, 2, prob["horizon_depth"])
    z0 = np.load(golden_case.HERE + "/golden/case_00.npz")   # same basin, same network
    net_prob, _ = golden_case.load("case_00")
    net = dict(net_prob["net"])
    net["rout_param"] = z["rout_param"]
    prob["net"] = net
    prob["inflowQ"] = np.zeros((0, n_days))
    ref = {"Qsim": np.stack([z[k] for k in z.files if k.startswith("Qsim/")]), "warming_days": warming,
           "final": {k[6:]: z[k] for k in z.files if k.startswith("final/")}}
    assert np.array_equal(z["net/L11_netPerm"], z0["net/L11_netPerm"])
    return prob, ref


def test_oracle_reproduces_penman_monteith_run():
    """MPR (aerodynamic + bulk surface resistance) -> PET case 3 -> cascade -> routing against the
    discharge of the reference's final calibration run: pins the FORCES constants cp0, rho0 and
    Psychro, which test_pet.pf fixes only to 1e-3.  gamma is known to 15 digits only."""
    import orc_mpr

    prob, ref = _case03(orc_mpr.init_lowres_level, orc_mpr.run_mpr)
    o = orc_run.OracleRun(prob)
    o.run(1, prob["time"]["nTimeSteps"])
    q = golden_case.daily_mean(o.mRM_runoff, ref["warming_days"])
    worst = parity.assert_close(q, ref["Qsim"], "case_03 daily discharge", rtol=1e-10, atol=0.0)
    for ours, theirs in (("L11_qTIN", "L11_qTIN"), ("L11_qTR", "L11_qTR"), ("L11_C1", "L11_C1"), ("L11_C2", "L11_C2")):
        parity.assert_close(o.R[ours], ref["final"][theirs], "case_03 " + theirs, rtol=1e-10)
    print("case_03: daily discharge max rel diff %.2e" % worst)
