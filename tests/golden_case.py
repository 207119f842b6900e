"""Load a golden fixture (tests/golden/case_*.npz, made by tests/golden/make_golden.py from the
reference's own saved outputs) as a problem dict in the mhm_b200.synth layout, plus the
reference's results to compare against."""
import datetime
import os

import numpy as np

from mhm_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load(case):
    z = np.load(os.path.join(HERE, "golden", case + ".npz"))
    soil_case, pet_case, rout_case = [int(x) for x in z["cases"]]
    P = {k[6:]: z[k] for k in z.files if k.startswith("param/")}
    n = P["L1_fSealed"].shape[-1]
    nH = P["L1_soilMoistSat"].shape[1]
    nLC = P["L1_fSealed"].shape[0]
    nLAI = P["L1_maxInter"].shape[1]
    ordinal0, n_days, warming = [int(x) for x in z["time"]]
    d0 = datetime.date.fromordinal(ordinal0)
    prob = {"nH": nH, "nLAI": nLAI, "nLC": nLC, "timestep_h": 1, "hourly": False,
            "soil_case": soil_case, "pet_case": pet_case, "rout_case": rout_case,
            "read_weights": False, "nCells": n, "nTstepForcingDay": 1}
    prob["processMatrix"] = synth.process_matrix(soil_case, pet_case, rout_case)
    lc = z["lc_years"]
    prob["time"] = {"jul_start": synth.JUL_1990_01_01 + (d0 - datetime.date(1990, 1, 1)).days,
                    "nTimeSteps": n_days * 24, "warming_days": warming, "timeStep_LAI_input": 0,
                    "lc_year_start": int(lc[0]), "LCyearId": np.asarray(lc[1:], dtype=np.int32)}
    # parameters the cascade does not use in this process selection still need storage
    full = synth.make_params(np.random.default_rng(0), n, nH, nLAI, nLC, pet_case)
    P["latitude"] = z["L1_lat"][None, None, :]
    for k, v in full.items():
        if k not in P:
            P[k] = v
        assert P[k].shape == v.shape, (k, P[k].shape, v.shape)
    prob["params"] = {k: np.ascontiguousarray(v) for k, v in P.items()}
    prob["forcing"] = {k[8:]: np.ascontiguousarray(z[k]) for k in z.files if k.startswith("forcing/")}
    prob["horizon_depth"] = np.ascontiguousarray(z["horizon_bnds"][:, 1])
    prob["states0"] = synth.default_states(n, nH, prob["horizon_depth"])
    outs = None
    if "out_flags" in z.files:
        # NetCDF name -> (outputFlxState index, 0-based horizon or -1)
        names = {"interception": (1, -1), "snowpack": (2, -1), "SM_Lall": (5, -1), "sealedSTW": (6, -1),
                 "unsatSTW": (7, -1), "satSTW": (8, -1), "PET": (9, -1), "aET": (10, -1), "Q": (11, -1),
                 "QD": (12, -1), "QIf": (13, -1), "QIs": (14, -1), "QB": (15, -1), "recharge": (16, -1),
                 "preEffect": (20, -1), "Qsm": (21, -1)}
        for h in range(nH):
            names["SWC_L%02d" % (h + 1)] = (3, h)
            names["SM_L%02d" % (h + 1)] = (4, h)
            names["soil_infil_L%02d" % (h + 1)] = (17, h)
            names["aET_L%02d" % (h + 1)] = (19, h)
        outs = {"flags": z["out_flags"], "timestep": int(z["out_timestep"]),
                "fields": {names[k[4:]]: z[k] for k in z.files if k.startswith("out/")},
                "time_bnds": z["out_time_bnds"]}
    if rout_case == 0:
        prob["net"] = None
        ref = {"final": {k[6:]: z[k] for k in z.files if k.startswith("final/")},
               "warming_days": warming, "n_days": n_days, "outputs": outs}
        return prob, ref
    nn = int(z["net/mask11"].sum())
    nLinks = int((z["net/L11_fromN"] > 0).sum())
    net = {"nNodes": nn, "nOutlets": nn - nLinks, "nCells1": n, "map_flag": 1,
           "fromN": z["net/L11_fromN"], "toN": z["net/L11_toN"], "netPerm": z["net/L11_netPerm"],
           "rOrder": z["net/L11_rOrder"], "L1_L11_Id": z["net/L1_L11_Id"],
           "L11_L1_Id": np.where(z["net/L11_L1_Id"] > 0, z["net/L11_L1_Id"], 1), "L1_areaCell": z["L1_areaCell_km2"],
           "L11_areaCell": z["net/L11_areaCell_km2"], "gaugeNodeList": z["net/gaugeNodeList"],
           "gaugeIndexList": np.arange(1, len(z["net/gaugeNodeList"]) + 1, dtype=np.int32),
           "nGaugesTotal": len(z["net/gaugeNodeList"]),
           "InflowGaugeNodeList": np.zeros(0, np.int32), "InflowGaugeIndexList": np.zeros(0, np.int32),
           "InflowGaugeHeadwater": np.zeros(0, np.int32), "nInflowTotal": 0, "processCase": rout_case,
           "L11_length": z["net/L11_length"], "L11_slope": z["net/L11_slope"],
           "L11_nLinkFracFPimp": z["net/L11_nLinkFracFPimp"], "rout_param": z["rout_param"],
           "TSrout": int(z["net/L11_TSrout"][0]) if "net/L11_TSrout" in z.files else 0,
           "celerity": float(z["celerity"][0])}
    if rout_case == 3:   # link locations on the L0 grid and the slope factor for L11_calc_celerity
        for k in ("fRow", "fCol", "tRow", "tCol"):
            net[k] = z["net/L11_" + k]
        net["slope_factor"] = float(z["slope_factor"][0])
    net = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in net.items()}
    prob["net"] = net
    prob["inflowQ"] = np.zeros((0, n_days))
    ref = {"final": {k[6:]: z[k] for k in z.files if k.startswith("final/")},
           "Qsim": np.stack([z[k] for k in z.files if k.startswith("Qsim/")]),
           "Qsim_text": z["Qsim_text"].T if "Qsim_text" in z.files else None, "warming_days": warming, "n_days": n_days, "outputs": outs}
    return prob, ref


def daily_mean(mRM_runoff, warming_days, nTstepDay=24):
    """mRM/mo_mrm_write.f90:142-150: mean of the model steps of each day after the warming"""
    q = np.asarray(mRM_runoff)[:, warming_days * nTstepDay:]
    return q.reshape(q.shape[0], -1, nTstepDay).sum(axis=2) / float(nTstepDay)


def load_mpr(case, init_lowres_level):
    """MPR problem (mhm_b200.synth_mpr layout) for the test basin with the gamma vector of a
    check case, and the L1 effective parameters the reference's MPR produced for it.
    init_lowres_level(mask0, cellsize0, target_resolution, cell_area0) -> grid dict: the
    oracle's or the library's restatement of common/mo_grid.f90:58-183."""
    z0 = np.load(os.path.join(HERE, "golden", "test_domain_l0.npz"))
    zc = np.load(os.path.join(HERE, "golden", case + ".npz"))
    soil_case, pet_case, _ = [int(x) for x in zc["cases"]]
    mask0 = z0["mask0"]
    n0 = int(mask0.sum())
    cs0 = float(z0["cellsize0"])
    mask1 = zc["mask1"]
    target = cs0 * mask0.shape[0] / mask1.shape[0]
    grid = init_lowres_level(mask0, cs0, target, np.full(n0, cs0 * cs0))
    db = {k[5:]: z0[k] for k in z0.files if k.startswith("soil/")}
    db["nSoil"], db["maxHor"] = int(db["nSoil"]), int(db["maxHor"])
    db["is_present"] = z0["is_present"]
    nH = zc["param/L1_soilMoistSat"].shape[1] if "param/L1_soilMoistSat" in zc.files else zc["horizon_bnds"].shape[0]
    hd = np.array(zc["horizon_bnds"][:, 1], dtype=np.float64)
    hd[nH - 1] = float(db.pop("HorizonDepth_last"))
    prob = {"nrows0": mask0.shape[1], "ncols0": mask0.shape[0], "nL0": n0,
            "mask0": np.ascontiguousarray(mask0, dtype=np.int32), "grid": grid, "nL1": grid["nCells1"],
            "nLC": z0["LCover0"].shape[0], "nLAI": 12, "nH": nH, "geoUnit0": z0["geoUnit0"],
            "soilId0": z0["soilId0"], "LCover0": z0["LCover0"], "Asp0": z0["Asp0"],
            "slope_emp0": z0["slope_emp0"], "y0": z0["y0"], "LAI0": z0["LAI0"], "soil_db": db,
            "HorizonDepth": hd, "GeoUnitList": z0["GeoUnitList"], "GeoUnitKar": z0["GeoUnitKar"],
            "fracSealed_CityArea": float(z0["fracSealed_CityArea"]), "param": zc["gamma"],
            "processMatrix": zc["processMatrix"], "soil_case": soil_case, "pet_case": pet_case}
    ref = {k[6:]: zc[k] for k in zc.files if k.startswith("param/")}
    ref["mask1"] = mask1
    if "L1_areaCell_km2" in zc.files:
        ref["L1_areaCell_km2"] = zc["L1_areaCell_km2"]
    return prob, ref
