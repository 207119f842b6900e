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
             xll=0.0, yll=0.0, LCover0=None, LCClassImp=2, routingCase=1):
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
    i.routingCase = routingCase
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
    res["streamNet0"] = np.zeros(n0, dtype=np.int32)
    o.streamNet0 = ip(res["streamNet0"])
    res["streamNet0"] = keep[-1]
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


def flow_accumulation(grid11, fDir11):
    """L11_flow_accumulation: L11_fAcc [km2] (nNodes) from L11_fDir and the cell areas [m2] of
    init_lowres_level's grid"""
    L = _lib.load()
    pi, pd = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.mrm_net_flow_accumulation.argtypes = [C.c_int32, C.c_int32, pi, pi, pd, pd]
    m11 = np.ascontiguousarray(grid11["mask1"], dtype=np.int32)
    fd = np.ascontiguousarray(fDir11, dtype=np.int32)
    area = np.ascontiguousarray(grid11["cellArea1"], dtype=np.float64)
    out = np.zeros(grid11["nCells1"])
    check(L.mrm_net_flow_accumulation(grid11["nrows1"], grid11["ncols1"], m11.ctypes.data_as(pi),
                                      fd.ctypes.data_as(pi), area.ctypes.data_as(pd), out.ctypes.data_as(pd)))
    return out


def calc_celerity(mask0, fDir0, slope0, net, slope_factor):
    """L11_calc_celerity on the outputs of net_init: (L11_celerity[nNodes], L0_celerity[nCells0])"""
    L = _lib.load()
    pi, pd = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.mrm_net_calc_celerity.argtypes = [C.c_int32, C.c_int32, pi, pi, pi, pd, C.c_int32, C.c_int32, pi, pi, pi, pi,
                                        pi, C.c_double, pd, pd]
    m0 = np.ascontiguousarray(mask0, dtype=np.int32)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    fd, sn, sl = i32(fDir0), i32(net["streamNet0"]), np.ascontiguousarray(slope0, dtype=np.float64)
    arrs = [i32(net[k]) for k in ("netPerm", "fRow", "fCol", "tRow", "tCol")]
    nn = len(arrs[0])
    c11, c0 = np.zeros(nn), np.zeros(int(m0.sum()))
    check(L.mrm_net_calc_celerity(m0.shape[1], m0.shape[0], m0.ctypes.data_as(pi), fd.ctypes.data_as(pi),
                                  sn.ctypes.data_as(pi), sl.ctypes.data_as(pd), nn, int(net["nLinks"]),
                                  *[a.ctypes.data_as(pi) for a in arrs], float(slope_factor),
                                  c11.ctypes.data_as(pd), c0.ctypes.data_as(pd)))
    return c11, c0


def update_param(length, celerity, nOutlets):
    """mrm_update_param for processCase(8) = 2 (scalar celerity) or 3 (L11_celerity): (C1, C2, TSrout)"""
    L = _lib.load()
    pd = C.POINTER(C.c_double)
    L.mrm_net_update_param.argtypes = [C.c_int32, C.c_int32, pd, pd, C.c_int32, pd, pd, pd]
    ln = np.ascontiguousarray(length, dtype=np.float64)
    cel = np.atleast_1d(np.ascontiguousarray(celerity, dtype=np.float64))
    c1, c2, ts = np.zeros(len(ln)), np.zeros(len(ln)), C.c_double(0.0)
    check(L.mrm_net_update_param(len(ln), int(nOutlets), ln.ctypes.data_as(pd), cel.ctypes.data_as(pd),
                                 0 if len(cel) == 1 else 1, c1.ctypes.data_as(pd), c2.ctypes.data_as(pd),
                                 C.byref(ts)))
    return c1, c2, ts.value
