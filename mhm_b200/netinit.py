"""ctypes wrapper of mrm_net_init (include/mhm_cuda.h, section N2): river-network initialisation
from the L0 flow direction / accumulation grids, host only."""
import ctypes as C

import numpy as np

from . import _cstruct, _lib
from ._lib import HEADER, check


class NetInputs(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mrm_net_inputs")


class NetOutputs(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mrm_net_outputs")


def net_init(mask0, fDir0, fAcc0, elev0, cellsize0, grid11, gaugeLoc0=None, gaugeIdList=(), coord_sys=0,
             xll=0.0, yll=0.0, LCover0=None, LCClassImp=2):
    """mask0: numpy bool (ncols0, nrows0) == Fortran (nrows0, ncols0); packed L0 vectors; grid11 =
    init_lowres_level(mask0, cellsize0, resolutionRouting) (mhm_b200.synth_mpr).  Returns a dict
    of the reference's L11_* network arrays."""
    L = _lib.load()
    L.mrm_net_init.argtypes = [C.POINTER(NetInputs), C.POINTER(NetOutputs)]
    L.mrm_net_init.restype = C.c_int
    keep = []

    def ip(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int32))

    def dp(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double))

    i = NetInputs()
    m0 = np.ascontiguousarray(mask0, dtype=np.int32)
    i.ncols0, i.nrows0 = m0.shape
    i.nrows11, i.ncols11, i.nNodes = grid11["nrows1"], grid11["ncols1"], grid11["nCells1"]
    i.nGauges, i.coord_sys, i.cellsize0, i.xllcorner0, i.yllcorner0 = len(gaugeIdList), coord_sys, cellsize0, xll, yll
    n0 = int(m0.sum())
    i.outlet_capacity = n0
    i.mask0, i.mask11 = ip(m0), ip(grid11["mask1"])
    i.fDir0, i.fAcc0, i.elev0 = ip(fDir0), ip(fAcc0), dp(elev0)
    if gaugeLoc0 is not None and len(gaugeIdList):
        i.gaugeLoc0, i.gaugeIdList = ip(gaugeLoc0), ip(np.asarray(gaugeIdList))
    for k in ("upper_bound", "lower_bound", "left_bound", "right_bound", "lowres_id_on_highres"):
        setattr(i, k, ip(grid11[k]))
    nn = i.nNodes
    o = NetOutputs()
    res = {}
    for k in ("fDir11", "rowOut", "colOut", "fromN", "toN", "rOrder", "netPerm", "fRow", "fCol", "tRow", "tCol"):
        res[k] = np.zeros(nn, dtype=np.int32)
        setattr(o, k, ip(res[k]))
        res[k] = keep[-1]
    res["gaugeNodeList"] = np.zeros(max(1, len(gaugeIdList)), dtype=np.int32)
    o.gaugeNodeList = ip(res["gaugeNodeList"])
    res["gaugeNodeList"] = keep[-1]
    for k in ("draSC0", "draCell0", "L0_rowOutlet", "L0_colOutlet"):
        res[k] = np.zeros(n0, dtype=np.int32)
        setattr(o, k, ip(res[k]))
        res[k] = keep[-1]
    for k in ("length", "slope"):
        res[k] = np.zeros(nn)
        setattr(o, k, dp(res[k]))
        res[k] = keep[-1]
    res["aFloodPlain"] = np.zeros(nn)
    o.aFloodPlain = dp(res["aFloodPlain"])
    res["aFloodPlain"] = keep[-1]
    res["floodPlain0"] = np.zeros(n0, dtype=np.int32)
    o.floodPlain0 = ip(res["floodPlain0"])
    res["floodPlain0"] = keep[-1]
    if LCover0 is not None:
        lc = np.ascontiguousarray(LCover0, dtype=np.int32)   # numpy (nLC, nCells0)
        i.LCover0, i.nLCoverScene, i.LCClassImp = ip(lc), lc.shape[0], LCClassImp
        res["nLinkFracFPimp"] = np.zeros((lc.shape[0], nn))
        o.nLinkFracFPimp = dp(res["nLinkFracFPimp"])
        res["nLinkFracFPimp"] = keep[-1]
    check(L.mrm_net_init(C.byref(i), C.byref(o)))
    res.update(nLinks=o.nLinks, nOutlets11=o.nOutlets11, L0_nOutlets=o.L0_nOutlets, nCells0=o.nCells0)
    res["gaugeNodeList"] = res["gaugeNodeList"][: len(gaugeIdList)]
    res["L0_rowOutlet"] = res["L0_rowOutlet"][: o.L0_nOutlets]
    res["L0_colOutlet"] = res["L0_colOutlet"][: o.L0_nOutlets]
    return res


def l1_l11_mapping(grid1, cellsize1, grid11, cellsize11):
    """L11_L1_mapping: (L1_L11_Id[nCells1], L11_L1_Id[nNodes]) from two init_lowres_level grids"""
    L = _lib.load()
    pi = C.POINTER(C.c_int32)
    L.mrm_net_l1_l11_mapping.argtypes = [C.c_int32, C.c_int32, pi, C.c_double, C.c_int32, C.c_int32, pi,
                                         C.c_double, pi, pi]
    m1 = np.ascontiguousarray(grid1["mask1"], dtype=np.int32)
    m11 = np.ascontiguousarray(grid11["mask1"], dtype=np.int32)
    a = np.zeros(grid1["nCells1"], dtype=np.int32)
    b = np.zeros(grid11["nCells1"], dtype=np.int32)
    check(L.mrm_net_l1_l11_mapping(grid1["nrows1"], grid1["ncols1"], m1.ctypes.data_as(pi), cellsize1,
                                   grid11["nrows1"], grid11["ncols1"], m11.ctypes.data_as(pi), cellsize11,
                                   a.ctypes.data_as(pi), b.ctypes.data_as(pi)))
    return a, b
