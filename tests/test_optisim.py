"""Calibration aggregates: the oracle's restatement of
mhm_interface_run_update_optisim (mo_mhm_interface_run.f90:745-861) + the BFI sums (:630-636)
against a literal numpy transcription driven by the oracle's own per-step history.  The
container optidata_sim is FORCES code (not vendored): parity unpinned for it, this test fixes
the semantics stated in oracle/mhm_oracle.h."""
import datetime

import numpy as np

import orc_run
from mhm_b200 import synth


def date_flags(prob):
    """(is_new_day, is_new_month, is_new_year, yId) AFTER the date increment of each step"""
    t = prob["time"]
    nT, dt_h = t["nTimeSteps"], prob["timestep_h"]
    start = datetime.datetime(*prob["start"])
    lc = np.asarray(t["LCyearId"])
    flags, yid = [], []
    y = int(lc[start.year - t["lc_year_start"]])
    for tt in range(1, nT + 1):
        a = start + datetime.timedelta(hours=(tt - 1) * dt_h)
        b = start + datetime.timedelta(hours=tt * dt_h)
        f = (a.date() != b.date(), (a.year, a.month) != (b.year, b.month), a.year != b.year)
        if f[2] and tt < nT:
            y = int(lc[b.year - t["lc_year_start"]])
        flags.append(f)
        yid.append(y)
    return flags, yid


def literal_optisim(prob, o, cfg, cell_area):
    """update_optisim + optidata_sim, written out step by step"""
    n, nH, nT = prob["nCells"], prob["nH"], prob["time"]["nTimeSteps"]
    warm = prob["time"]["warming_days"] * (24 // prob["timestep_h"])
    flags, yid = date_flags(prob)
    P = prob["params"]
    data = {k: np.zeros((v[1], n)) for k, v in cfg.items()}
    ts_, cnt = {k: 1 for k in cfg}, {k: 0 for k in cfg}
    qbf = qt = 0.0

    def flag(k, tt):
        return flags[tt - 1][-cfg[k][0] - 1]

    def add(k, v):
        if ts_[k] <= cfg[k][1]:
            data[k][ts_[k] - 1] = data[k][ts_[k] - 1] + v

    def average(k):
        if ts_[k] <= cfg[k][1]:
            with np.errstate(invalid="ignore", divide="ignore"):
                data[k][ts_[k] - 1] = data[k][ts_[k] - 1] / float(cnt[k])
        ts_[k] += 1
        cnt[k] = 0

    for tt in range(1, nT + 1):
        if tt - warm <= 0:
            continue
        sm = o.hist("L1_soilMoist", tt)
        y = yid[tt - 1] - 1
        qbf += (o.hist("L1_baseflow", tt) * cell_area).sum() / n
        qt += (o.hist("L1_total_runoff", tt) * cell_area).sum() / n
        if "sm" in cfg:
            if flag("sm", tt):
                average("sm")
            if tt != nT:
                nh = cfg["sm"][2]
                a, b = np.zeros(n), np.zeros(n)
                for h in range(nh):
                    a = a + sm[h]
                for h in range(nh):
                    b = b + P["L1_soilMoistSat"][y, h]
                add("sm", a / b)
                cnt["sm"] += 1
        if "et" in cfg:
            if flag("et", tt):
                ts_["et"] += 1
            if tt != nT:
                fS = P["L1_fSealed"][y, 0]
                a = np.zeros(n)
                for h in range(nH):
                    a = a + o.hist("L1_aETSoil", tt)[h]
                add("et", a * (1.0 - fS) + o.hist("L1_aETCanopy", tt) + o.hist("L1_aETSealed", tt) * fS)
        if "tws" in cfg:
            if flag("tws", tt):
                average("tws")
            if tt != nT:
                add("tws", o.hist("L1_inter", tt) + o.hist("L1_snowPack", tt) + o.hist("L1_sealSTW", tt)
                    + o.hist("L1_unsatSTW", tt) + o.hist("L1_satSTW", tt))
                cnt["tws"] += 1
                for h in range(nH):
                    add("tws", sm[h])
    return data, qbf, qt


def n_windows(prob, timeStepInput):
    flags, _ = date_flags(prob)
    warm = prob["time"]["warming_days"] * (24 // prob["timestep_h"])
    return sum(1 for tt in range(warm + 1, prob["time"]["nTimeSteps"] + 1) if flags[tt - 1][-timeStepInput - 1])


def make(hourly=True, n_days=40, warming=2, nH=3, start=(1990, 12, 5)):
    prob = synth.make_problem(nx=9, ny=6, n_days=n_days, nH=nH, hourly=hourly, soil_case=1,
                              pet_case=-1 if hourly else 0, routing=False, start=start)
    prob["time"]["warming_days"] = warming
    prob["start"] = start
    return prob


def test_oracle_optisim_equals_literal_transcription():
    prob = make()
    nT, n = prob["time"]["nTimeSteps"], prob["nCells"]
    cfg = {"sm": (-1, n_windows(prob, -1), 2), "et": (-2, n_windows(prob, -2) + 1), "tws": (-2, n_windows(prob, -2))}
    area = np.random.default_rng(5).uniform(0.5, 2.0, n)
    o = orc_run.OracleRun(prob, history=True, optisim=cfg, bfi=True, cell_area=area)
    o.run(1, 300)
    o.run(301, nT)  # any split of the time axis
    data, qbf, qt = literal_optisim(prob, o, cfg, area)
    for k in cfg:
        np.testing.assert_array_equal(o.opt[k], data[k])
        assert np.isfinite(o.opt[k]).all() and (o.opt[k][: cfg[k][1] - 1] != 0).all()
    # windows: a daily slot closes BEFORE the value of the day's last step is added -> the first
    # slot holds nTstepDay - 1 values, the last value of the run is never added
    assert o.d.opt_avg_ts[0] == cfg["sm"][1] + 1 and o.d.opt_avg_cnt[0] == 0
    np.testing.assert_allclose(o.d.bfi_qBF_sum, qbf, rtol=1e-14)
    np.testing.assert_allclose(o.d.bfi_qT_sum, qt, rtol=1e-14)
    assert qt > qbf > 0
