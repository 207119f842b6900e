"""Known-answer tests transcribed from the reference's pFUnit suite
(/root/reference/src/tests/*.pf), run against the CPU oracle.
Tolerance 1e-3 absolute, as upstream (`real(dp) :: t = 0.001_dp`)."""
import ctypes as C
import math

import numpy as np
import pytest

import orc

T = 1e-3
EPS = np.finfo(np.float64).eps


def close(ref, res):
    if isinstance(ref, float) and math.isnan(ref):
        return math.isnan(res)
    return abs(ref - res) <= T


# --- test_canopy_interc.pf:27-50 -------------------------------------------------
@pytest.mark.parametrize(
    "args,interc0,ref",
    [
        ((0.1, 0.45, 0.39), 0.345, (0.35, 0.285, 0.1)),
        ((0.1, 0.45, -0.39), 0.345, (0.0, 0.0, -0.045)),
        ((0.1, -0.45, 0.39), 0.345, (0.0, 1.185, -0.45)),
    ],
)
def test_canopy_interc(args, interc0, ref):
    L = orc.lib()
    interc, thr, evap = orc.ref(interc0), orc.ref(), orc.ref()
    L.orc_canopy_interc(*args, C.byref(interc), C.byref(thr), C.byref(evap))
    assert close(ref[0], interc.value) and close(ref[1], thr.value) and close(ref[2], evap.value)


# --- test_snow_accum_melt.pf:32-91 ------------------------------------------------
@pytest.mark.parametrize(
    "args,pack0,ref",
    [
        # (snow_pack, deg_day, melt, prec_effect, rain, snow)
        ((0.5, 0.232, 0.078, 0.208, 8.31, 1.0, 0.195), 0.0, (0.0, 0.182, 0.0, 0.195, 0.195, 0.0)),
        ((0.5, 0.232, 0.078, 0.208, -8.31, 1.0, 0.195), 0.0, (0.195, 0.182, 0.0, 0.0, 0.0, 0.195)),
        ((0.5, 0.232, 0.078, 0.208, 3.0, 1.0, 0.195), 0.3, (0.0, 0.182, 0.3, 0.495, 0.195, 0.0)),
        ((0.5, 0.232, 0.232, 0.208, 3.0, 1.0, 0.195), 0.6, (0.136, 0.232, 0.464, 0.659, 0.195, 0.0)),
    ],
)
def test_snow_accum_melt(args, pack0, ref):
    L = orc.lib()
    out = [orc.ref(pack0)] + [orc.ref() for _ in range(5)]
    L.orc_snow_accum_melt(*args, *[C.byref(o) for o in out])
    for r, o in zip(ref, out):
        assert close(r, o.value)


# --- test_soil_moisture.pf:42-133 --------------------------------------------------
def _soil(case, fs, wts, pet, ec, sat, fr, fc, wp, ex, jc1, aetc, pe, ro0, st0, sm0):
    L = orc.lib()
    nH = len(sat)
    a = lambda x: np.ascontiguousarray(np.array(x, dtype=np.float64))
    sat, fr, fc, wp, ex, sm = map(a, (sat, fr, fc, wp, ex, sm0))
    inf, aet = np.zeros(nH), np.zeros(nH)
    ro, st, aets = orc.ref(ro0), orc.ref(st0), orc.ref()
    L.orc_soil_moisture(case, fs, wts, pet, ec, nH, 1, orc.dptr(sat), orc.dptr(fr), orc.dptr(fc),
                        orc.dptr(wp), orc.dptr(ex), jc1, aetc, pe, C.byref(ro), C.byref(st),
                        orc.dptr(inf), orc.dptr(sm), orc.dptr(aet), C.byref(aets))
    return ro.value, st.value, inf, sm, aet, aets.value


def test_soil_moisture_set1():
    ro, st, inf, sm, aet, aets = _soil(2, 0.03, 0.5, 0.3, 1.0, [91.0, 65.0], [0.6, 0.6], [67.0, 67.0],
                                      [21.2, 21.2], [2.2, 2.2], 0.5, 0.0, 0.0, 0.0, 0.0, [43.0, 70.0])
    assert close(0.0, ro) and close(0.0, st) and close(0.0, aets)
    assert np.allclose(inf, [0.0, 0.0], atol=T, rtol=0)
    assert np.allclose(sm, [42.887, 69.887], atol=T, rtol=0)
    assert np.allclose(aet, [0.112, 0.112], atol=T, rtol=0)


def test_soil_moisture_set2():
    ro, st, inf, sm, aet, aets = _soil(1, 0.03, 0.0, 0.3, 1.0, [1.0], [0.6], [0.0], [5.0], [2.2], 0.5,
                                      0.5, 10.0, 0.0, 0.5, [0.0])
    assert close(10.5, ro) and close(0.0, st) and close(0.0, aets)
    assert close(9.0, inf[0]) and close(1.0, sm[0]) and close(0.0, aet[0])


def test_soil_moisture_set3():
    ro, st, inf, sm, aet, aets = _soil(1, 0.03, 0.5, 0.3, 1.0, [0.0], [0.6], [0.0], [5.0], [2.2], 0.5,
                                      0.5, 1.0, 0.0, 1.0, [0.5 * EPS])
    assert close(1.5, ro) and close(0.5, st) and close(0.0, aets)
    assert close(1.0, inf[0]) and close(EPS, sm[0]) and close(0.0, aet[0])


# --- test_soil_moisture.pf:136-167 --------------------------------------------------
def test_feddes():
    L = orc.lib()
    assert close(0.0, L.orc_feddes_et_reduction(38.5, 77.0, 43.6, 0.6))
    assert close(0.04, L.orc_feddes_et_reduction(35.5, 71.1, 33.0, 0.618))
    assert close(0.618, L.orc_feddes_et_reduction(35.5, 35.5, 33.0, 0.618))


def test_jarvis():
    L = orc.lib()
    assert close(0.315, L.orc_jarvis_et_reduction(48.1, 92.6, 32.2, 0.6, 0.5))
    assert close(0.0, L.orc_jarvis_et_reduction(128.4, 319.0, 144.5, 0.4, 0.5))
    assert close(0.6, L.orc_jarvis_et_reduction(91.3, 50.5, 38.2, 0.6, 0.5))


# --- test_runoff.pf:32-106 -----------------------------------------------------------
@pytest.mark.parametrize(
    "pefec,unsat0,ref",
    [
        # (sat, unsat, slow, fast, perc)
        (0.0, 2.27, (35.304, 0.0, 2.27, 0.0, 0.0)),
        (20.0, 2.27, (35.304, 0.0, 6.862, 15.408, 0.0)),
        (0.0, -2.27, (32.055, 0.979, 0.0, 0.0, -3.249)),
    ],
)
def test_runoff_unsat_zone(pefec, unsat0, ref):
    L = orc.lib()
    sat, unsat, slow, fast, perc = orc.ref(35.304), orc.ref(unsat0), orc.ref(), orc.ref(), orc.ref()
    L.orc_runoff_unsat_zone(1.2417, 1.4312, 3.547, 0.285, 1.0, pefec, 17.926, C.byref(sat),
                            C.byref(unsat), C.byref(slow), C.byref(fast), C.byref(perc))
    for r, o in zip(ref, (sat, unsat, slow, fast, perc)):
        assert close(r, o.value), (ref, [x.value for x in (sat, unsat, slow, fast, perc)])


def test_runoff_sat_zone():
    L = orc.lib()
    sat, bf = orc.ref(17.0), orc.ref()
    L.orc_runoff_sat_zone(0.000417, C.byref(sat), C.byref(bf))
    assert close(16.992, sat.value) and close(0.007, bf.value)
    sat = orc.ref(-17.0)
    L.orc_runoff_sat_zone(0.000417, C.byref(sat), C.byref(bf))
    assert close(0.0, sat.value) and close(0.0, bf.value)


def test_total_runoff():
    L = orc.lib()
    tr = orc.ref()
    L.orc_L1_total_runoff(0.0284, 0.0, 5.0, 5.365, 0.0, C.byref(tr))
    assert close(10.07, tr.value)


# --- test_pet.pf:19-90 ----------------------------------------------------------------
def test_pet():
    L = orc.lib()
    assert close(1349.437, L.orc_pet_hargreaves(0.8, 17.8, 12.0, 20.0, 8.0, 23.1637, 150))
    assert close(4.978, L.orc_pet_priestly(1.26, 200.0, 10.0))
    assert close(0.0, L.orc_pet_priestly(1.26, 0.0, 10.0))
    assert close(1.074, L.orc_pet_penman(200.0, 10.0, 1.7, 60.0, 70.0, 1.0, 1.0))
    assert close(16.34, L.orc_extraterr_rad_approx(150, math.pi / 180.0 * 23.1637))
    assert close(0.082, L.orc_slope_satpressure(10.0))
    assert close(1.228, L.orc_sat_vap_pressure(10.0))


# --- test_meteo_temporal_tools.pf:26-62 -------------------------------------------------
def test_temporal_disagg():
    L = orc.lib()
    for v, r in zip((2.74, 3.32), (0.822, 0.996)):
        assert close(r, L.orc_temporal_disagg_meteo_weights(v, 0.3, 0.0))
    for v, r in zip((2.74, 3.32), (0.472, 0.646)):
        assert close(r, L.orc_temporal_disagg_meteo_weights(v, 0.3, 0.5))
    for (day, nts), r in zip(((1, 24.0), (0, 24.0), (1, 1.0)), (0.2055, 0.0228, 2.74)):
        assert close(r, L.orc_temporal_disagg_flux_daynight(day, nts, 2.74, 0.9, 0.1))
    cases = ((1, 24.0, 0), (0, 24.0, 0), (1, 24.0, 1), (0, 24.0, 1), (1, 1.0, 0))
    for (day, nts, add), r in zip(cases, (4.932, 0.548, 3.64, 2.84, 2.74)):
        assert close(r, L.orc_temporal_disagg_state_daynight(day, nts, 2.74, 0.9, 0.1, add))


# --- calendar: julday/caldat against Python's proleptic Gregorian calendar ---------------
def test_calendar():
    import datetime

    L = orc.lib()
    base = datetime.date(1990, 1, 1)
    assert L.orc_julday(1, 1, 1990) == 2447893  # JDN of 1990-01-01 (noon)
    dd, mm, yy = C.c_int32(), C.c_int32(), C.c_int32()
    for off in range(-20000, 30000, 37):
        dt = base + datetime.timedelta(days=off)
        j = L.orc_julday(dt.day, dt.month, dt.year)
        assert j == 2447893 + off
        L.orc_caldat(j, C.byref(dd), C.byref(mm), C.byref(yy))
        assert (dd.value, mm.value, yy.value) == (dt.day, dt.month, dt.year)
        assert L.orc_doy(dt.day, dt.month, dt.year) == dt.timetuple().tm_yday
