"""Synthetic L0 morphology, soil data base and global parameters for the MPR path.

Mirrors what the reference holds after mpr_initialize (MPR/mo_mpr_startup.f90): packed L0
fields, the L0 -> L1 remap of init_lowres_level, soilDB with depth weights Wd
(MPR/mo_soil_database.f90:401-490 restated here for input generation only) and the flat
parameter vector with processMatrix(:, 2:3) slicing it (MPR/mo_mpr_read_config.f90:390-988).
Default parameter values are those of the reference's mhm_parameter.nml.
"""
import ctypes as C

import numpy as np

from . import _lib
from .interface import _pd, _pi
from ._lib import check

DEFAULTS = {
    "interception1": [0.15],
    "snow1": [1.0, 1.5, 0.5, 0.5, 0.5, 3.0, 3.5, 4.0],
    "soilmoisture1": [3.4, 0.1, 0.6, 0.76, 0.0009, -0.264, 0.89, -0.001, -0.324, -0.585, 0.0125, 0.0063,
                      60.960, 0.97, 0.93, 0.02, 1.75],
    "directRunoff1": [0.5],
    "PETminus1": [0.3, 0.8, 1.3, 1.5, -0.7],
    "PET0": [0.9, 0.1, 180.0],
    "PET1": [0.93, 0.19, 171.0, 0.0023],
    "PET2": [1.19, 0.058],
    "PET3": [15.0, 0.02, 0.11, 0.64, 0.095, 0.075, 56.0],
    "interflow1": [85.0, 7.0, 1.5, 15.0, 0.125],
    "percolation1": [35.0, -1.0, 1.0],
    "routing1": [0.325, 0.075, 2.0, 0.1, 0.3],
}
DEFAULTS["soilmoisture2"] = DEFAULTS["soilmoisture1"] + [0.5]
DEFAULTS["soilmoisture3"] = DEFAULTS["soilmoisture1"][:13] + [0.975, 0.975, 0.975, 1.75, 0.09, 0.98, 0.15, 0.25, 0.5]
DEFAULTS["soilmoisture4"] = DEFAULTS["soilmoisture3"][:-1]


def global_parameters(soil_case=1, pet_case=-1, n_geo=10, rng=None, jitter=0.0):
    """flat gamma vector + processMatrix (numpy (3, 11)); jitter scales every value by a random
    factor in [1-jitter, 1+jitter] (ensemble members)"""
    pet = {-1: "PETminus1", 0: "PET0", 1: "PET1", 2: "PET2", 3: "PET3"}[pet_case]
    blocks = [("interception1", 1), ("snow1", 1), ("soilmoisture%d" % soil_case, soil_case),
              ("directRunoff1", 1), (pet, pet_case), ("interflow1", 1), ("percolation1", 1),
              ("routing1", 1)]
    pm = np.zeros((3, 11), dtype=np.int32)
    vals, end = [], 0
    for p, (name, case) in enumerate(blocks):
        v = list(DEFAULTS[name])
        vals += v
        end += len(v)
        pm[:, p] = [case, len(v), end]
    geo = list(np.linspace(100.0, 1000.0, n_geo)) if rng is None else list(rng.uniform(50.0, 1000.0, n_geo))
    vals += geo
    end += n_geo
    pm[:, 8] = [1, n_geo, end]
    pm[:, 9] = [0, 0, end]
    pm[:, 10] = [0, 0, end]
    g = np.array(vals, dtype=np.float64)
    if rng is not None and jitter > 0:
        f = rng.uniform(1 - jitter, 1 + jitter, len(g))
        keep = np.zeros(len(g), dtype=bool)
        # aspect threshold, Ks curve slope and karstic gain stay fixed like in the namelist (flag 0)
        for k, v in enumerate(g):
            keep[k] = v in (60.960, 180.0, 171.0)
        g = np.where(keep, g, g * f)
    return g, pm


def init_lowres_level(mask0, cellsize0, target_resolution, cell_area0=None):
    """mhm_grid_init_lowres_level (host helper of the library): L0 -> L1 maps.
    mask0: numpy bool (ncols0, nrows0) == Fortran (nrows0, ncols0)."""
    L = _lib.load()
    m0 = np.ascontiguousarray(mask0, dtype=np.int32)
    ncols0, nrows0 = m0.shape
    nr, nc, n1 = C.c_int32(), C.c_int32(), C.c_int32()
    null_i, null_d = C.POINTER(C.c_int32)(), C.POINTER(C.c_double)()
    ca = null_d if cell_area0 is None else _pd(np.ascontiguousarray(cell_area0, dtype=np.float64))
    check(L.mhm_grid_init_lowres_level(nrows0, ncols0, _pi(m0), ca, cellsize0, target_resolution,
                                       C.byref(nr), C.byref(nc), C.byref(n1), null_i, null_i, null_d,
                                       null_i, null_i, null_i, null_i, null_i, null_i))
    n = n1.value
    out = {"nrows1": nr.value, "ncols1": nc.value, "nCells1": n,
           "mask1": np.zeros((nc.value, nr.value), dtype=np.int32),
           "cellCoor": np.zeros((2, n), dtype=np.int32), "cellArea1": np.zeros(n),
           "upper_bound": np.zeros(n, dtype=np.int32), "lower_bound": np.zeros(n, dtype=np.int32),
           "left_bound": np.zeros(n, dtype=np.int32), "right_bound": np.zeros(n, dtype=np.int32),
           "n_subcells": np.zeros(n, dtype=np.int32),
           "lowres_id_on_highres": np.zeros((ncols0, nrows0), dtype=np.int32)}
    check(L.mhm_grid_init_lowres_level(nrows0, ncols0, _pi(m0), ca, cellsize0, target_resolution,
                                       C.byref(nr), C.byref(nc), C.byref(n1), _pi(out["mask1"]),
                                       _pi(out["cellCoor"]), _pd(out["cellArea1"]), _pi(out["upper_bound"]),
                                       _pi(out["lower_bound"]), _pi(out["left_bound"]),
                                       _pi(out["right_bound"]), _pi(out["n_subcells"]),
                                       _pi(out["lowres_id_on_highres"])))
    return out


def make_soil_db(rng, n_soil, horizon_depth, tillage_depth=200.0, max_hor=5):
    """random soil data base + depth weights (mo_soil_database.f90:401-490, iFlag_soilDB = 0)"""
    nH = len(horizon_depth)
    nHor = rng.integers(2, max_hor + 1, n_soil).astype(np.int32)
    mh = int(nHor.max())
    UD = np.zeros((mh, n_soil))
    LD = np.zeros((mh, n_soil))
    sand = np.zeros((mh, n_soil))
    clay = np.zeros((mh, n_soil))
    DbM = np.zeros((mh, n_soil))
    nTill = np.zeros(n_soil, dtype=np.int32)
    RZ = np.zeros(n_soil)
    for s in range(n_soil):
        # first horizon ends at the tillage depth, deeper ones are 100..900 mm thick
        if nHor[s] >= 3 and rng.random() < 0.5:  # two tillage horizons
            bounds = [0.0, 0.5 * tillage_depth, tillage_depth] + list(
                tillage_depth + np.cumsum(rng.integers(1, 10, nHor[s] - 2) * 100.0))
        else:
            bounds = [0.0, tillage_depth] + list(tillage_depth + np.cumsum(rng.integers(1, 10, nHor[s] - 1) * 100.0))
        # the data base must reach below the deepest fixed model horizon
        bounds[-1] = max(bounds[-1], horizon_depth[nH - 2] + 100.0 * rng.integers(1, 8))
        for j in range(nHor[s]):
            UD[j, s], LD[j, s] = bounds[j], bounds[j + 1]
            sd = rng.uniform(5.0, 90.0)
            sand[j, s] = sd
            clay[j, s] = rng.uniform(2.0, max(3.0, 95.0 - sd))
            DbM[j, s] = rng.uniform(1.1, 1.8)
        nTill[s] = sum(1 for j in range(nHor[s]) if UD[j, s] < tillage_depth)
        RZ[s] = np.rint(LD[nHor[s] - 1, s])
    Wd = np.zeros((mh, nH, n_soil))
    acc = 0.5
    hd = np.array(horizon_depth, dtype=np.float64)
    for s in range(n_soil):
        Wd[nHor[s]:, :, s] = -9999.0
        hd[nH - 1] = RZ[s]
        for jj in range(nH):
            f = 0.0 if jj == 0 else hd[jj - 1]
            t = hd[jj] - acc
            lf = lt = -1
            for kk in range(nHor[s]):
                if UD[kk, s] <= f <= LD[kk, s] - acc:
                    lf = kk
                if UD[kk, s] <= t <= LD[kk, s] - acc:
                    lt = kk
            assert lf >= 0 and lt >= lf, (s, jj, lf, lt)
            if lf == lt:
                Wd[lf, jj, s] = 1.0
            else:
                Wd[lf, jj, s] = LD[lf, s] - f
                Wd[lt, jj, s] = (t + acc) - UD[lt, s]
                for kk in range(lf + 1, lt):
                    Wd[kk, jj, s] = LD[kk, s] - UD[kk, s]
                div = hd[jj] if jj == 0 else hd[jj] - hd[jj - 1]
                Wd[: nHor[s], jj, s] = Wd[: nHor[s], jj, s] / div
    return {"nSoil": n_soil, "maxHor": mh, "is_present": np.ones(n_soil, dtype=np.int32), "nHorizons": nHor,
            "nTillHorizons": nTill, "sand": sand, "clay": clay, "DbM": DbM, "Wd": Wd, "RZdepth": RZ}


def make_mpr_problem(nx0=60, ny0=40, factor=5, nLC=2, nLAI=12, nH=2, soil_case=1, pet_case=-1, n_soil=30,
                     n_geo=10, fill=0.8, seed=20261017):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ny0, 0:nx0]  # numpy (ncols0, nrows0): first Fortran index = x
    r = ((xx - nx0 / 2.0) / (nx0 / 2.0)) ** 2 + ((yy - ny0 / 2.0) / (ny0 / 2.0)) ** 2 + rng.normal(0, 0.08, xx.shape)
    mask0 = r <= np.quantile(r, fill)
    n0 = int(mask0.sum())
    grid = init_lowres_level(mask0, 100.0, 100.0 * factor, np.full(n0, 1.0e4))
    horizon_depth = np.array([200.0 * (i + 1) for i in range(nH)])
    horizon_depth[nH - 1] = 0.0  # the last model horizon reaches the soil type's root-zone depth
    db = make_soil_db(rng, n_soil, horizon_depth)
    gamma, pm = global_parameters(soil_case, pet_case, n_geo)
    geo_list = np.arange(1, n_geo + 1, dtype=np.int32) * 3  # not contiguous: exercises minloc
    lai_class = rng.uniform(0.3, 6.0, (nLAI, 3))
    LC = rng.integers(1, 4, (nLC, n0)).astype(np.int32)
    prob = {
        "nrows0": nx0, "ncols0": ny0, "nL0": n0, "mask0": np.ascontiguousarray(mask0, dtype=np.int32),
        "grid": grid, "nL1": grid["nCells1"], "nLC": nLC, "nLAI": nLAI, "nH": nH,
        "geoUnit0": rng.choice(geo_list, n0).astype(np.int32),
        "soilId0": rng.integers(1, n_soil + 1, n0).astype(np.int32),
        "LCover0": LC,
        "Asp0": rng.uniform(0.0, 360.0, n0),
        "slope_emp0": rng.uniform(0.001, 1.0, n0),
        "y0": np.where(rng.random(n0) < 0.9, rng.uniform(30.0, 60.0, n0), rng.uniform(-40.0, -10.0, n0)),
        "LAI0": np.ascontiguousarray(lai_class[:, LC[0] - 1] * rng.uniform(0.9, 1.1, (nLAI, n0))),
        "soil_db": db, "HorizonDepth": horizon_depth,
        "GeoUnitList": geo_list, "GeoUnitKar": (rng.random(n_geo) < 0.3).astype(np.int32),
        "fracSealed_CityArea": 0.6, "param": gamma, "processMatrix": pm,
        "soil_case": soil_case, "pet_case": pet_case,
    }
    return prob


def shard_mpr_problem(prob, cells):
    """The MPR inputs of ONE shard of a domain cut by L1 cells (sub-catchment sharding, BASELINE
    config 4): the L0 cells under the shard's L1 cells, cropped to their bounding box, with the
    shard's rows of the L0 -> L1 remap.  An L1 cell's parameters depend only on the L0 cells of its
    own rectangle (MPR/mo_mpr_eval.f90:120-154, mo_upscaling_operators.f90), so a shard's MPR needs
    no exchange and equals the whole domain's MPR on its cells bit for bit.
    cells: ascending 0-based L1 cell indices."""
    cells = np.asarray(cells, dtype=np.int64)
    g = prob["grid"]
    ub, lb = np.asarray(g["upper_bound"])[cells], np.asarray(g["lower_bound"])[cells]   # first index (rows), 1-based
    le, ri = np.asarray(g["left_bound"])[cells], np.asarray(g["right_bound"])[cells]    # second index (columns)
    r0, r1, c0, c1 = int(ub.min()), int(lb.max()), int(le.min()), int(ri.max())
    mask_full = np.asarray(prob["mask0"]) != 0          # numpy (ncols0, nrows0) == Fortran (nrows0, ncols0)
    sel = np.zeros_like(mask_full)
    for k in range(len(cells)):                         # union of the shard's rectangles
        sel[le[k] - 1: ri[k], ub[k] - 1: lb[k]] = True
    sel &= mask_full
    packed = sel[mask_full]                             # the shard's entries of every packed L0 vector
    sub = dict(prob)
    sub["nrows0"], sub["ncols0"] = r1 - r0 + 1, c1 - c0 + 1
    sub["mask0"] = np.ascontiguousarray(sel[c0 - 1: c1, r0 - 1: r1], dtype=np.int32)
    sub["nL0"] = int(packed.sum())
    sub["nL1"] = len(cells)
    sub["grid"] = {"upper_bound": (ub - (r0 - 1)).astype(np.int32), "lower_bound": (lb - (r0 - 1)).astype(np.int32),
                   "left_bound": (le - (c0 - 1)).astype(np.int32), "right_bound": (ri - (c0 - 1)).astype(np.int32),
                   "n_subcells": np.asarray(g["n_subcells"])[cells].astype(np.int32), "nCells1": len(cells)}
    for k in ("geoUnit0", "soilId0", "Asp0", "slope_emp0", "y0"):
        sub[k] = np.ascontiguousarray(np.asarray(prob[k])[packed])
    for k in ("LCover0", "LAI0"):
        sub[k] = np.ascontiguousarray(np.asarray(prob[k])[:, packed])
    # the one global number MPR needs: the soil type of the whole domain's last L0 cell (see
    # mpr_l0_inputs.lastSoilId0)
    sub["lastSoilId0"] = int(prob.get("lastSoilId0", 0)) or int(np.asarray(prob["soilId0"])[-1])
    return sub


def set_mpr_inputs(dom, prob):
    """mpr_cuda_set_l0 + mpr_cuda_set_soildb for an interface.Domain"""
    L, g, db = dom.L, prob["grid"], prob["soil_db"]
    keep = []

    def ip(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return _pi(a)

    def dp(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return _pd(a)

    l0 = _lib.MprL0Inputs()
    l0.nrows0, l0.ncols0 = prob["nrows0"], prob["ncols0"]
    l0.mask0 = ip(prob["mask0"])
    for k in ("upper_bound", "lower_bound", "left_bound", "right_bound", "n_subcells"):
        setattr(l0, k, ip(g[k]))
    l0.geoUnit0, l0.soilId0, l0.LCover0 = ip(prob["geoUnit0"]), ip(prob["soilId0"]), ip(prob["LCover0"])
    l0.Asp0, l0.slope_emp0, l0.y0 = dp(prob["Asp0"]), dp(prob["slope_emp0"]), dp(prob["y0"])
    l0.gridded_LAI0 = dp(prob["LAI0"])
    l0.lastSoilId0 = int(prob.get("lastSoilId0", 0))
    check(L.mpr_cuda_set_l0(dom.h, dom.id, C.byref(l0)))
    sd = _lib.MprSoilDb()
    sd.nSoilTypes, sd.maxHorizons, sd.nGeoUnits = db["nSoil"], db["maxHor"], len(prob["GeoUnitList"])
    sd.is_present, sd.nHorizons, sd.nTillHorizons = ip(db["is_present"]), ip(db["nHorizons"]), ip(db["nTillHorizons"])
    sd.sand, sd.clay, sd.DbM, sd.Wd, sd.RZdepth = dp(db["sand"]), dp(db["clay"]), dp(db["DbM"]), dp(db["Wd"]), dp(db["RZdepth"])
    sd.HorizonDepth_mHM = dp(prob["HorizonDepth"])
    sd.GeoUnitList, sd.GeoUnitKar = ip(prob["GeoUnitList"]), ip(prob["GeoUnitKar"])
    sd.fracSealed_CityArea = prob["fracSealed_CityArea"]
    check(L.mpr_cuda_set_soildb(dom.h, dom.id, C.byref(sd)))
    del keep


def mpr_eval(dom, param, member=0):
    p = np.ascontiguousarray(param, dtype=np.float64)
    check(dom.L.mpr_cuda_eval(dom.h, dom.id, member, _pd(p), len(p)))


# (dim2, dim3) of every array MPR writes, as functions of (nH, nLAI, nLC)
MPR_OUTPUTS = {
    "L1_fSealed": lambda h, a, c: (1, c), "L1_alpha": lambda h, a, c: (1, c),
    "L1_degDayInc": lambda h, a, c: (1, c), "L1_degDayMax": lambda h, a, c: (1, c),
    "L1_degDayNoPre": lambda h, a, c: (1, c), "L1_fRoots": lambda h, a, c: (h, c),
    "L1_maxInter": lambda h, a, c: (a, 1), "L1_karstLoss": lambda h, a, c: (1, 1),
    "L1_kFastFlow": lambda h, a, c: (1, c), "L1_kSlowFlow": lambda h, a, c: (1, c),
    "L1_kBaseFlow": lambda h, a, c: (1, c), "L1_kPerco": lambda h, a, c: (1, c),
    "L1_soilMoistFC": lambda h, a, c: (h, c), "L1_soilMoistSat": lambda h, a, c: (h, c),
    "L1_soilMoistExp": lambda h, a, c: (h, c), "L1_jarvis_thresh_c1": lambda h, a, c: (1, 1),
    "L1_tempThresh": lambda h, a, c: (1, c), "L1_unsatThresh": lambda h, a, c: (1, 1),
    "L1_sealedThresh": lambda h, a, c: (1, 1), "L1_wiltingPoint": lambda h, a, c: (h, c),
    "L1_petLAIcorFactor": lambda h, a, c: (a, c), "L1_fAsp": lambda h, a, c: (1, 1),
    "L1_HarSamCoeff": lambda h, a, c: (1, 1), "L1_PrieTayAlpha": lambda h, a, c: (a, 1),
    "L1_aeroResist": lambda h, a, c: (a, c), "L1_surfResist": lambda h, a, c: (a, 1),
}


def outputs_for(soil_case, pet_case):
    names = [k for k in MPR_OUTPUTS if k not in ("L1_petLAIcorFactor", "L1_fAsp", "L1_HarSamCoeff",
                                                  "L1_PrieTayAlpha", "L1_aeroResist", "L1_surfResist",
                                                  "L1_jarvis_thresh_c1")]
    if soil_case in (2, 3):
        names.append("L1_jarvis_thresh_c1")
    names += {-1: ["L1_petLAIcorFactor"], 0: ["L1_fAsp"], 1: ["L1_fAsp", "L1_HarSamCoeff"],
              2: ["L1_PrieTayAlpha"], 3: ["L1_aeroResist", "L1_surfResist"]}[pet_case]
    return names
