"""Run the MPR part of the CPU oracle on a synth_mpr problem.  Test infrastructure."""
import ctypes as C
import os

import numpy as np

import orc
from mhm_b200 import _cstruct, synth_mpr

HDR = os.path.join(orc.ODIR, "mpr_oracle.h")


class L0Grid(C.Structure):
    _fields_ = _cstruct.parse_struct(HDR, "orc_l0_grid")


class MprIn(C.Structure):
    _fields_ = _cstruct.parse_struct(HDR, "orc_mpr_in")


class MprOut(C.Structure):
    _fields_ = _cstruct.parse_struct(HDR, "orc_mpr_out")


_OUT_FIELD = {
    "L1_fSealed": "fSealed", "L1_alpha": "alpha", "L1_degDayInc": "degDayInc", "L1_degDayMax": "degDayMax",
    "L1_degDayNoPre": "degDayNoPre", "L1_fAsp": "fAsp", "L1_HarSamCoeff": "HarSamCoeff",
    "L1_PrieTayAlpha": "PrieTayAlpha", "L1_aeroResist": "aeroResist", "L1_surfResist": "surfResist",
    "L1_fRoots": "fRoots", "L1_kFastFlow": "kFastFlow", "L1_kSlowFlow": "kSlowFlow",
    "L1_kBaseFlow": "kBaseFlow", "L1_kPerco": "kPerco", "L1_karstLoss": "karstLoss",
    "L1_soilMoistFC": "soilMoistFC", "L1_soilMoistSat": "soilMoistSat", "L1_soilMoistExp": "soilMoistExp",
    "L1_jarvis_thresh_c1": "jarvis_thresh_c1", "L1_tempThresh": "tempThresh",
    "L1_unsatThresh": "unsatThresh", "L1_sealedThresh": "sealedThresh", "L1_wiltingPoint": "wiltingPoint",
    "L1_maxInter": "maxInter", "L1_petLAIcorFactor": "petLAIcorFactor",
}


def _lib():
    L = orc.lib()
    if not getattr(L, "_mpr_ready", False):
        pd, pi, i, d = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32, C.c_double
        L.orc_upscale_arithmetic_mean.argtypes = [C.POINTER(L0Grid), pd, pd]
        L.orc_upscale_harmonic_mean.argtypes = [C.POINTER(L0Grid), pd, pd]
        L.orc_upscale_geometric_mean.argtypes = [C.POINTER(L0Grid), d, pd, pd]
        L.orc_L0_fractionalCover_in_Lx.argtypes = [C.POINTER(L0Grid), pi, i, pd]
        for f in ("orc_upscale_arithmetic_mean", "orc_upscale_harmonic_mean", "orc_upscale_geometric_mean",
                  "orc_L0_fractionalCover_in_Lx"):
            getattr(L, f).restype = None
        L.orc_mpr.argtypes = [C.POINTER(MprIn), C.POINTER(MprOut)]
        L.orc_mpr.restype = i
        L.orc_init_lowres_level.argtypes = [i, i, pi, pd, d, d, i, i, pi, pi, pd, pi, pi, pi, pi, pi, pi]
        L.orc_init_lowres_level.restype = i
        L.orc_calculate_grid_properties.argtypes = [i, i, d, d, d, d, pi, pi, pd, pd, pd]
        L.orc_calculate_grid_properties.restype = None
        L._mpr_ready = True
    return L


def grid_struct(prob, keep):
    g = L0Grid()
    gr = prob["grid"]
    g.nrows0, g.ncols0, g.nL1 = prob["nrows0"], prob["ncols0"], prob["nL1"]
    for fld, arr in (("mask0", prob["mask0"]), ("upper", gr["upper_bound"]), ("lower", gr["lower_bound"]),
                     ("left", gr["left_bound"]), ("right", gr["right_bound"]), ("nsub", gr["n_subcells"])):
        a = np.ascontiguousarray(arr, dtype=np.int32)
        keep.append(a)
        setattr(g, fld, orc.iptr(a))
    return g


def upscale(prob, op, x, class_id=0, nodata=-9999.0):
    L, keep = _lib(), []
    g = grid_struct(prob, keep)
    out = np.zeros(prob["nL1"])
    if op == "frac":
        xi = np.ascontiguousarray(x, dtype=np.int32)
        L.orc_L0_fractionalCover_in_Lx(C.byref(g), orc.iptr(xi), class_id, orc.dptr(out))
    else:
        xd = np.ascontiguousarray(x, dtype=np.float64)
        if op == "arith":
            L.orc_upscale_arithmetic_mean(C.byref(g), orc.dptr(xd), orc.dptr(out))
        elif op == "harm":
            L.orc_upscale_harmonic_mean(C.byref(g), orc.dptr(xd), orc.dptr(out))
        else:
            L.orc_upscale_geometric_mean(C.byref(g), nodata, orc.dptr(xd), orc.dptr(out))
    return out


def run_mpr(prob, param=None):
    """returns dict reference-name -> numpy (dim3, dim2, nL1)"""
    L, keep = _lib(), []
    n1, nH, nLAI, nLC = prob["nL1"], prob["nH"], prob["nLAI"], prob["nLC"]
    db, gr = prob["soil_db"], prob["grid"]

    def ip(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return orc.iptr(a)

    def dp(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return orc.dptr(a)

    m = MprIn()
    m.nrows0, m.ncols0, m.nL0, m.nL1 = prob["nrows0"], prob["ncols0"], prob["nL0"], n1
    m.nLC, m.nLAI, m.nH = nLC, nLAI, nH
    m.nSoil, m.maxHor, m.nGeo = db["nSoil"], db["maxHor"], len(prob["GeoUnitList"])
    pm = np.ascontiguousarray(prob["processMatrix"], dtype=np.int32)
    m.nProc = pm.shape[1]
    p = prob["param"] if param is None else param
    m.nParam = len(p)
    m.mask0 = ip(prob["mask0"])
    m.upper, m.lower, m.left, m.right = ip(gr["upper_bound"]), ip(gr["lower_bound"]), ip(gr["left_bound"]), ip(gr["right_bound"])
    m.nsub = ip(gr["n_subcells"])
    m.geoUnit0, m.soilId0, m.LCover0 = ip(prob["geoUnit0"]), ip(prob["soilId0"]), ip(prob["LCover0"])
    m.Asp0, m.slope_emp0, m.y0, m.LAI0 = dp(prob["Asp0"]), dp(prob["slope_emp0"]), dp(prob["y0"]), dp(prob["LAI0"])
    m.is_present, m.nHorizons, m.nTillHorizons = ip(db["is_present"]), ip(db["nHorizons"]), ip(db["nTillHorizons"])
    m.sand, m.clay, m.DbM, m.Wd, m.RZdepth = dp(db["sand"]), dp(db["clay"]), dp(db["DbM"]), dp(db["Wd"]), dp(db["RZdepth"])
    m.HorizonDepth = dp(prob["HorizonDepth"])
    m.GeoUnitList, m.GeoUnitKar = ip(prob["GeoUnitList"]), ip(prob["GeoUnitKar"])
    m.fracSealed_CityArea = prob["fracSealed_CityArea"]
    m.processMatrix = ip(pm)
    m.param = dp(p)
    m.lastSoilId0 = int(prob.get("lastSoilId0", 0))
    o = MprOut()
    out = {}
    for name, fld in _OUT_FIELD.items():
        d2, d3 = synth_mpr.MPR_OUTPUTS[name](nH, nLAI, nLC)
        out[name] = np.zeros((d3, d2, n1))
        setattr(o, fld, orc.dptr(out[name]))
    rc = L.orc_mpr(C.byref(m), C.byref(o))
    assert rc == 0
    return out


def init_lowres_level(mask0, cellsize0, target_resolution, cell_area0=None):
    """the oracle's init_lowres_level with the same result dict as synth_mpr.init_lowres_level"""
    L = _lib()
    m0 = np.ascontiguousarray(mask0, dtype=np.int32)
    ncols0, nrows0 = m0.shape
    nr, nc = C.c_int32(), C.c_int32()
    xll, yll, cs = C.c_double(), C.c_double(), C.c_double()
    L.orc_calculate_grid_properties(nrows0, ncols0, 0.0, 0.0, cellsize0, target_resolution, C.byref(nr),
                                    C.byref(nc), C.byref(xll), C.byref(yll), C.byref(cs))
    n0 = int(m0.sum())
    area = np.full(n0, cellsize0 * cellsize0) if cell_area0 is None else np.ascontiguousarray(cell_area0, dtype=np.float64)
    nmax = nr.value * nc.value
    mask1 = np.zeros(nmax, dtype=np.int32)
    coor = np.zeros(2 * nmax, dtype=np.int32)
    a1 = np.zeros(nmax)
    up, lo, le, ri, ns = (np.zeros(nmax, dtype=np.int32) for _ in range(5))
    ids = np.zeros(nrows0 * ncols0, dtype=np.int32)
    n1 = L.orc_init_lowres_level(nrows0, ncols0, orc.iptr(m0), orc.dptr(area), cellsize0, target_resolution,
                                 nr.value, nc.value, orc.iptr(mask1), orc.iptr(coor), orc.dptr(a1), orc.iptr(up),
                                 orc.iptr(lo), orc.iptr(le), orc.iptr(ri), orc.iptr(ns), orc.iptr(ids))
    return {"nrows1": nr.value, "ncols1": nc.value, "nCells1": n1,
            "mask1": mask1.reshape(nc.value, nr.value), "cellArea1": a1[:n1].copy(),
            "upper_bound": up[:n1].copy(), "lower_bound": lo[:n1].copy(), "left_bound": le[:n1].copy(),
            "right_bound": ri[:n1].copy(), "n_subcells": ns[:n1].copy(),
            "lowres_id_on_highres": ids.reshape(ncols0, nrows0)}
