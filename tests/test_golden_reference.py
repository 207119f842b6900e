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

CASES = ["case_00", "case_02", "case_09", "case_10", "case_12",
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
        parity.assert_bit_exact(o.R[ours], ref["final"][theirs], "%s %s" % (case, theirs))
    q = golden_case.daily_mean(o.mRM_runoff, ref["warming_days"])
    # the daily mean is a sum of 24 values: summation order of the Fortran `sum` intrinsic
    # (gfortran may vectorise it) is the only freedom left -> 4 ulp
    worst = parity.assert_close(q, ref["Qsim"], case + " daily discharge", rtol=1e-15, atol=0.0)
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
                parity.assert_close(dom.get_routing_state(ours), ref["final"][theirs],
                                    "%s %s (%s)" % (case, theirs, mode), rtol=parity.RTOL_Q)
            q = golden_case.daily_mean(dom.get_runoff(), ref["warming_days"])
            wq = parity.assert_close(q, ref["Qsim"], case + " daily discharge", rtol=parity.RTOL_Q)
            msg += ", daily discharge %.2e" % wq
        print(msg)
