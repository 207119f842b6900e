"""Calibration aggregates and BFI sums on the device: mhm_cuda_set_optisim +
run_steps against the oracle's restatement of mhm_interface_run_update_optisim
(mo_mhm_interface_run.f90:745-861) and of the BFI sums (:630-636)."""
import numpy as np
import pytest

import orc_run
import parity
from mhm_b200 import driver, interface, synth
from test_optisim import make, n_windows

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("mode", ["strict", "fast"])
@pytest.mark.parametrize("hourly,ts_sm,ts_et,ts_tws,nhor,warming", [(True, -1, -1, -2, 2, 2), (False, -2, -3, -1, 3, 0),
                                                                    (True, -3, -2, -1, 1, 30)])
def test_optisim_equals_oracle(ctx, mode, hourly, ts_sm, ts_et, ts_tws, nhor, warming):
    """daily / monthly / yearly slots over a year and land-cover scene change, warming period,
    run issued in four calls, gridded outputs switched on at the same time"""
    prob = make(hourly=hourly, n_days=40, warming=warming, nH=3)
    nT, n = prob["time"]["nTimeSteps"], prob["nCells"]
    nw = lambda ts: max(1, n_windows(prob, ts))  # a yearly slot may never close inside the evaluation period
    cfg = {"sm": (ts_sm, nw(ts_sm), nhor), "et": (ts_et, nw(ts_et)), "tws": (ts_tws, nw(ts_tws))}
    area = np.random.default_rng(5).uniform(0.5, 2.0, n)
    flags = np.zeros(21, dtype=np.int32)
    flags[[4, 9]] = 1
    o = orc_run.OracleRun(prob, optisim=cfg, bfi=True, cell_area=area, outputs=(flags, -2))
    o.run(1, nT)
    ctx.set_math_mode(mode)
    dom = fresh(ctx, prob)
    dom.set_optisim(sm=cfg["sm"], et=cfg["et"], tws=cfg["tws"], bfi=True)
    dom.set_outputs(flags, -2)
    cuts = [1, 24, 131, 600, nT + 1] if hourly else [1, 2, 7, 30, nT + 1]
    for a, b in zip(cuts[:-1], cuts[1:]):
        dom.run_steps(a, b - a)
    worst = 0.0
    for k in cfg:
        got = dom.get_optisim(k)
        assert np.isfinite(got).all() and np.abs(got).max() > 0
        worst = max(worst, parity.assert_close(got, o.opt[k], "optisim " + k))
    qbf, qt = dom.get_bfi_sums(area)
    np.testing.assert_allclose([qbf, qt], [o.d.bfi_qBF_sum, o.d.bfi_qT_sum], rtol=1e-9)
    assert dom.output_windows()  # the last call closed the run's final output window
    print("%s: aggregates max rel diff %.2e, BFI %.6f vs %.6f" % (mode, worst, qbf / qt, o.d.bfi_qBF_sum / o.d.bfi_qT_sum))
    ctx.set_math_mode("strict")


def test_optisim_members_with_routing(ctx):
    """two ensemble members, routing on (fused node runoff must stay correct with the OUT kernel)"""
    prob = synth.make_problem(nx=12, ny=8, n_days=6, hourly=True, routing=True, start=(1991, 1, 29))
    prob["start"] = (1991, 1, 29)
    prob["time"]["warming_days"] = 1
    nT, n = prob["time"]["nTimeSteps"], prob["nCells"]
    rng = np.random.default_rng(3)
    members = [prob["params"], {k: (v * rng.uniform(0.95, 1.05) if k in ("L1_kPerco", "L1_alpha") else v)
                                for k, v in prob["params"].items()}]
    cfg = {"sm": (-1, n_windows(prob, -1), 2), "tws": (-2, n_windows(prob, -2))}
    refs = []
    for P in members:
        o = orc_run.OracleRun(prob, params=P, optisim=cfg, bfi=True)
        o.run(1, nT)
        refs.append(o)
    ctx.set_math_mode("fast")
    dom = fresh(ctx, prob, nMembers=2, member_params=members)
    dom.set_optisim(sm=cfg["sm"], tws=cfg["tws"], bfi=True)
    dom.run_steps(1, 50)
    dom.run_steps(51, nT - 50)
    for m, o in enumerate(refs):
        for k in cfg:
            parity.assert_close(dom.get_optisim(k, member=m), o.opt[k], "member %d %s" % (m, k))
        qbf, qt = dom.get_bfi_sums(prob["net"]["L1_areaCell"], member=m)
        np.testing.assert_allclose([qbf, qt], [o.d.bfi_qBF_sum, o.d.bfi_qT_sum], rtol=1e-9)
        parity.assert_close(dom.get_runoff(member=m), o.mRM_runoff, "gauge discharge", rtol=parity.RTOL_Q)
    with pytest.raises(interface._lib.MhmCudaError):
        dom.set_optisim(sm=(-1, 3, 9))  # more horizons than the model has (mo_mhm_read_config.f90:189)
    with pytest.raises(interface._lib.MhmCudaError):
        dom.set_optisim(et=(5, 3))
    ctx.set_math_mode("strict")
