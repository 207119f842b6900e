"""ctypes loader for the CPU oracle (oracle/libmhm_oracle.so).

Test infrastructure: imported only from tests/, bench.py's CPU arms and smoke().
The orc_domain struct is mirrored by parsing oracle/mhm_oracle.h, so the two
cannot drift apart.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "libmhm_oracle.so")

_CT = {
    "int32_t": C.c_int32,
    "int64_t": C.c_int64,
    "double": C.c_double,
}


def build(force=False):
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs
    ):
        subprocess.check_call(["make", "-C", ODIR, "-s"])
    return LIB


def parse_struct(header, name):
    """Return ctypes fields for `typedef struct name { ... } name;` in header."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), txt, re.S)
    assert m, name
    fields = []
    for line in m.group(1).split(";"):
        line = line.strip()
        if not line:
            continue
        mm = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(\w+)(\[(\d+)\])?$", line)
        assert mm, line
        base = _CT[mm.group(2)]
        if mm.group(3):
            ct = C.POINTER(base)
        elif mm.group(6):
            ct = base * int(mm.group(6))
        else:
            ct = base
        fields.append((mm.group(4), ct))
    return fields


class OrcDomain(C.Structure):
    _fields_ = parse_struct(os.path.join(ODIR, "mhm_oracle.h"), "orc_domain")


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        d = C.c_double
        i = C.c_int32
        pd = C.POINTER(C.c_double)
        pi = C.POINTER(C.c_int32)
        L.orc_canopy_interc.argtypes = [d, d, d, pd, pd, pd]
        L.orc_canopy_interc.restype = None
        L.orc_snow_accum_melt.argtypes = [d] * 7 + [pd] * 6
        L.orc_snow_accum_melt.restype = None
        L.orc_feddes_et_reduction.argtypes = [d] * 4
        L.orc_feddes_et_reduction.restype = d
        L.orc_jarvis_et_reduction.argtypes = [d] * 5
        L.orc_jarvis_et_reduction.restype = d
        L.orc_soil_moisture.argtypes = (
            [i, d, d, d, d, i, C.c_int64] + [pd] * 5 + [d, d, d] + [pd] * 6
        )
        L.orc_soil_moisture.restype = None
        L.orc_runoff_unsat_zone.argtypes = [d] * 7 + [pd] * 5
        L.orc_runoff_unsat_zone.restype = None
        L.orc_runoff_sat_zone.argtypes = [d, pd, pd]
        L.orc_runoff_sat_zone.restype = None
        L.orc_L1_total_runoff.argtypes = [d] * 5 + [pd]
        L.orc_L1_total_runoff.restype = None
        L.orc_pet_hargreaves.argtypes = [d] * 6 + [i]
        L.orc_pet_hargreaves.restype = d
        L.orc_pet_priestly.argtypes = [d] * 3
        L.orc_pet_priestly.restype = d
        L.orc_pet_penman.argtypes = [d] * 7
        L.orc_pet_penman.restype = d
        L.orc_extraterr_rad_approx.argtypes = [i, d]
        L.orc_extraterr_rad_approx.restype = d
        L.orc_slope_satpressure.argtypes = [d]
        L.orc_slope_satpressure.restype = d
        L.orc_sat_vap_pressure.argtypes = [d]
        L.orc_sat_vap_pressure.restype = d
        L.orc_temporal_disagg_meteo_weights.argtypes = [d, d, d]
        L.orc_temporal_disagg_meteo_weights.restype = d
        L.orc_temporal_disagg_flux_daynight.argtypes = [i, d, d, d, d]
        L.orc_temporal_disagg_flux_daynight.restype = d
        L.orc_temporal_disagg_state_daynight.argtypes = [i, d, d, d, d, i]
        L.orc_temporal_disagg_state_daynight.restype = d
        L.orc_julday.argtypes = [i, i, i]
        L.orc_julday.restype = i
        L.orc_caldat.argtypes = [i, pi, pi, pi]
        L.orc_caldat.restype = None
        L.orc_doy.argtypes = [i, i, i]
        L.orc_doy.restype = i
        L.orc_L11_runoff_acc.argtypes = [i, i, pd, pd, pi, pd, pi, i, i, pd]
        L.orc_L11_runoff_acc.restype = None
        L.orc_add_inflow.argtypes = [i, pi, pi, pi, pd, pd]
        L.orc_add_inflow.restype = None
        L.orc_L11_routing.argtypes = [i, i, pi, pi, pi, pd, pd, pd, i, pi, pi, pd, pd, pd]
        L.orc_L11_routing.restype = None
        L.orc_reg_rout.argtypes = [pd, i, i, pd, pd, pd, d, pd, pd]
        L.orc_reg_rout.restype = None
        L.orc_mrm_update_param_case2.argtypes = [i, i, pd, d, pd, pd]
        L.orc_mrm_update_param_case2.restype = d
        L.orc_stream_net.argtypes = [i, i, pi, i, pi, pi, pi, pi, pi, pi]
        L.orc_stream_net.restype = None
        L.orc_length_floor.argtypes = [i, pd]
        L.orc_length_floor.restype = None
        L.orc_calc_celerity.argtypes = [i, i, pi, pi, pi, pd, i, i, pi, pi, pi, pi, pi, d, pd]
        L.orc_calc_celerity.restype = None
        L.orc_mrm_update_param_case3.argtypes = [i, i, pd, pd, pd, pd]
        L.orc_mrm_update_param_case3.restype = d
        L.orc_flow_accumulation.argtypes = [i, i, pi, pi, pd, pd]
        L.orc_flow_accumulation.restype = None
        L.orc_flux_record_size.argtypes = [i]
        L.orc_flux_record_size.restype = i
        L.orc_run.argtypes = [C.POINTER(OrcDomain), i, i]
        L.orc_run.restype = i
        L.orc_time_indices.argtypes = [C.POINTER(OrcDomain), i] + [pi] * 8
        L.orc_time_indices.restype = None
        L.orc_routing_order_linear.argtypes = [i, i, pi, pi, pi, pi]
        L.orc_routing_order_linear.restype = i
        _lib = L
    return _lib


def routing_order(nNodes, fromN, toN, nLinks=None):
    """L11_routing_order through the oracle's linear-time restatement: (rOrder, netPerm), padded to
    nNodes like mhm_b200.interface.routing_order (which this replaces for the CPU arms of bench.py)"""
    fromN = np.ascontiguousarray(fromN, dtype=np.int32)
    toN = np.ascontiguousarray(toN, dtype=np.int32)
    nLinks = len(fromN) if nLinks is None else nLinks
    rOrder = np.full(nNodes, -9999, dtype=np.int32)
    netPerm = np.full(nNodes, -9999, dtype=np.int32)
    rc = lib().orc_routing_order_linear(nNodes, nLinks, iptr(fromN), iptr(toN), iptr(rOrder), iptr(netPerm))
    if rc != 0:
        raise ValueError("orc_routing_order_linear: %s" % ("the link graph has a cycle" if rc == 1 else "out of memory"))
    return rOrder, netPerm


def dptr(a):
    if a is None:
        return C.POINTER(C.c_double)()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    if a is None:
        return C.POINTER(C.c_int32)()
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def ref(x=0.0):
    return C.c_double(x)
