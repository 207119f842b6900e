"""CPU-only tests of the library's host logic: the C ABI loads and exports every declared
symbol, the calendar matches the oracle's restatement of datetimeinfo, and the O(N) routing
order equals a literal transcription of L11_routing_order."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
import orc_run
from mhm_b200 import _cstruct, _lib, interface, synth


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    names = _cstruct.declared_functions(_lib.HEADER)
    assert len(names) > 40
    for n in names:
        assert hasattr(L, n), "libmhm_cuda.so does not export %s" % n
    assert b"sm_100a" in L.mhm_cuda_version()


def test_fortran_module_binds_every_declared_symbol():
    """mhm_b200/fortran/mo_mhm_cuda.F90 carries one bind(C) interface per entry point of the
    header, made public (mhm_cuda_last_error stays private behind mhm_cuda_check)"""
    import re

    src = open(os.path.join(os.path.dirname(_lib.HEADER), "..", "mhm_b200", "fortran", "mo_mhm_cuda.F90")).read()
    bound = set(re.findall(r"bind\(C,\s*name\s*=\s*'(\w+)'\)", src))
    public = set()
    for stmt in re.findall(r"^\s*public\s*::((?:.*&\s*\n)*.*)$", src, flags=re.M):
        public |= set(re.findall(r"\w+", stmt))
    names = _cstruct.declared_functions(_lib.HEADER)
    assert len(names) > 40
    for n in names:
        assert n in bound, "mo_mhm_cuda.F90 has no bind(C) interface for %s" % n
        assert n in public or n == "mhm_cuda_last_error", "%s is not public in mo_mhm_cuda.F90" % n
    assert bound <= set(names), "bind(C) names unknown to the header: %s" % sorted(bound - set(names))
    # same number of arguments on both sides
    hdr = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER).read(), flags=re.S)
    joined = re.sub(r"&\s*\n\s*", "", src)
    for n in names:
        c = re.search(r"\b%s\s*\(([^;{}]*?)\)\s*;" % n, hdr).group(1).strip()
        f = re.search(r"function\s+%s\s*\(([^)]*)\)" % n, joined).group(1).strip()
        nc = 0 if c in ("", "void") else c.count(",") + 1
        nf = 0 if f == "" else f.count(",") + 1
        assert nc == nf, "%s: %d arguments in the header, %d in mo_mhm_cuda.F90" % (n, nc, nf)


def test_fortran_derived_types_match_header_structs():
    """every `type, bind(C)` of mo_mhm_cuda.F90 has the header struct's components: same names,
    same order, same size in bytes"""
    import re

    src = open(os.path.join(os.path.dirname(_lib.HEADER), "..", "mhm_b200", "fortran", "mo_mhm_cuda.F90")).read()
    src = re.sub(r"&\s*\n\s*", "", src)
    kind_bytes = {"c_int8_t": 1, "c_int16_t": 2, "c_int32_t": 4, "c_int": 4, "c_int64_t": 8, "c_double": 8}
    types = re.findall(r"type, bind\(C\) :: (\w+)\n(.*?)end type", src, flags=re.S)
    assert len(types) >= 8
    for name, body in types:
        comps = []
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" not in line:
                continue
            decl, rest = line.split("::", 1)
            if "c_ptr" in decl:
                size = C.sizeof(C.c_void_p)
            else:
                size = kind_bytes[re.search(r"\((\w+)\)", decl).group(1)]
            dim = re.search(r"dimension\((\d+)\)", decl)
            for c in re.split(r",(?![^(]*\))", rest):
                c = c.split("=")[0].strip()
                if not c:
                    continue
                own = re.search(r"\((\d+)\)", c)
                n = int(own.group(1)) if own else int(dim.group(1)) if dim else 1
                comps.append((re.sub(r"\(.*\)", "", c).lower(), size * n))
        fields = [(f.lower(), C.sizeof(t)) for f, t in _cstruct.parse_struct(_lib.HEADER, name)]
        assert comps == fields, "%s differs between mo_mhm_cuda.F90 and the header" % name


def test_no_gpu_means_error_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MhmCudaError, match="no CUDA device|no CPU fallback"):
        interface.Context()


@pytest.mark.parametrize("lai_mode", [0, -1, -2, -3])
@pytest.mark.parametrize("timestep_h,hourly", [(1, True), (1, False), (2, False), (24, False)])
def test_calendar_matches_oracle(lai_mode, timestep_h, hourly):
    prob = synth.make_problem(nx=4, ny=3, n_days=800, routing=False, timestep_h=timestep_h,
                              hourly=hourly, start=(1991, 11, 17), lc_switch_year=1993,
                              timeStep_LAI_input=lai_mode)
    prob["forcing"] = {k: v[:1] for k, v in prob["forcing"].items()}  # not needed here
    n = prob["time"]["nTimeSteps"]
    o = orc_run.OracleRun(prob).time_indices(n)
    idx = interface.time_indices(prob["time"], timestep_h, prob["nTstepForcingDay"], 1, n)
    for key in ("month", "hour", "yId", "iLAI", "iMeteoTS", "isday", "doy", "year"):
        got = np.array([getattr(s, key) for s in idx])
        assert np.array_equal(got, o[key]), key
    # a later window equals the tail of the full sequence
    idx2 = interface.time_indices(prob["time"], timestep_h, prob["nTstepForcingDay"], n - 49, 50)
    assert [s.iLAI for s in idx2] == [s.iLAI for s in idx][-50:]
    assert o["yId"].min() == 1 and o["yId"].max() == 2


def literal_routing_order(nNodes, fromN, toN):
    """mRM/mo_mrm_net_startup.f90:779-842 transcribed loop for loop (O(nLinks^2))."""
    nLinks = len(fromN)
    r = np.ones(nLinks, dtype=np.int64)
    for ii in range(nLinks):
        for jj in range(nLinks):
            if jj == ii:
                continue
            if fromN[ii] == toN[jj]:
                r[ii] = -9
            if r[ii] == -9:
                break
    kk = 0
    for ii in range(nLinks):
        if r[ii] == 1:
            kk += 1
            r[ii] = kk
    while nLinks and r.min() < 0:
        for ii in range(nLinks):
            if r[ii] != -9:
                continue
            flag = True
            for jj in range(nLinks):
                if jj == ii or fromN[ii] != toN[jj]:
                    continue
                elif not (fromN[ii] == toN[jj] and r[jj] > 0):
                    flag = False
                    break
            if flag:
                kk += 1
                r[ii] = kk
    perm = np.zeros(nLinks, dtype=np.int64)
    for ii in range(nLinks):
        perm[r[ii] - 1] = ii + 1
    return r, perm


def random_forest(rng, n):
    """random multi-outlet forest with arbitrary node numbering; links ascending in fromN"""
    parent = np.zeros(n + 1, dtype=np.int64)
    order = rng.permutation(n) + 1
    for pos, node in enumerate(order):
        if pos == 0 or rng.random() < 0.08:
            parent[node] = 0
        else:
            # prefer recent nodes so that chains get deep
            lo = max(0, pos - int(rng.integers(1, 12)))
            parent[node] = order[int(rng.integers(lo, pos))]
    fromN = np.array([k for k in range(1, n + 1) if parent[k] > 0], dtype=np.int32)
    toN = np.array([parent[k] for k in fromN], dtype=np.int32)
    return fromN, toN


def test_routing_order_equals_literal_reference():
    rng = np.random.default_rng(7)
    for trial in range(300):
        n = int(rng.integers(2, 60))
        fromN, toN = random_forest(rng, n)
        if len(fromN) == 0:
            continue
        r_ref, p_ref = literal_routing_order(n, fromN, toN)
        rOrder, netPerm = interface.routing_order(n, fromN, toN)
        nl = len(fromN)
        assert np.array_equal(rOrder[:nl], r_ref), trial
        assert np.array_equal(netPerm[:nl], p_ref), trial


def test_oracle_linear_routing_order_equals_literal_reference():
    """the oracle's own linear-time L11_routing_order (used by bench.py's CPU arms, which must not
    touch the product library) against the loop-for-loop transcription, and against the library"""
    import orc

    rng = np.random.default_rng(23)
    for trial in range(200):
        n = int(rng.integers(2, 60))
        fromN, toN = random_forest(rng, n)
        r_ref, p_ref = literal_routing_order(n, fromN, toN)
        rOrder, netPerm = orc.routing_order(n, fromN, toN)
        assert np.array_equal(rOrder[: len(fromN)], r_ref), trial
        assert np.array_equal(netPerm[: len(fromN)], p_ref), trial
    net = synth.scheidegger_network(np.random.default_rng(2), 60, 40)
    a = orc.routing_order(net["nNodes"], net["fromN"], net["toN"])
    b = interface.routing_order(net["nNodes"], net["fromN"], net["toN"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    with pytest.raises(ValueError):
        orc.routing_order(3, np.array([1, 2, 3], dtype=np.int32), np.array([2, 3, 1], dtype=np.int32))


def test_routing_order_scheidegger():
    rng = np.random.default_rng(3)
    net = synth.scheidegger_network(rng, 14, 9)
    r_ref, p_ref = literal_routing_order(net["nNodes"], net["fromN"], net["toN"])
    rOrder, netPerm = interface.routing_order(net["nNodes"], net["fromN"], net["toN"])
    nl = len(net["fromN"])
    assert np.array_equal(netPerm[:nl], p_ref)
    assert net["nOutlets"] >= 1


def test_routing_order_rejects_cycle():
    with pytest.raises(_lib.MhmCudaError, match="cycle"):
        interface.routing_order(3, np.array([1, 2, 3], dtype=np.int32), np.array([2, 3, 1], dtype=np.int32))
