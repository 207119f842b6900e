"""CPU tests of the MPR side: init_lowres_level against the reference's own known-answer test
(src/tests/test_grid.pf:16-97) for both the oracle and the library's host helper, upscaling
operators on hand-checked fields, and sanity of the oracle's full MPR."""
import ctypes as C

import numpy as np
import pytest

import orc
import orc_mpr
from mhm_b200 import synth_mpr


def grid_kat_mask():
    """highres mask of test_grid.pf:22-38 as numpy (ncols, nrows) == Fortran (nrows=6, ncols=7)"""
    v1 = [1, 0, 0, 0, 1, 1, 0]
    v2 = [0, 1, 1, 0, 1, 1, 1]
    m = np.zeros((7, 6), dtype=np.int32)  # [j][i]
    for i in range(3):
        for j in range(7):
            m[j, 2 * i] = v1[j]
            m[j, 2 * i + 1] = v2[j]
    return m


REF = {"lower": [5, 6, 5, 6], "upper": [1, 6, 1, 6], "left": [1, 1, 6, 6], "right": [5, 5, 7, 7],
       "nsub": [12, 3, 7, 2], "area": [300.0, 75.0, 175.0, 50.0], "coor": [[1, 2, 1, 2], [1, 1, 2, 2]]}


def test_init_lowres_level_kat_library_helper():
    m = grid_kat_mask()
    n0 = int(m.sum())
    g = synth_mpr.init_lowres_level(m, 10.0, 50.0, np.full(n0, 25.0))
    assert (g["nrows1"], g["ncols1"], g["nCells1"]) == (2, 2, 4)
    assert g["lower_bound"].tolist() == REF["lower"] and g["upper_bound"].tolist() == REF["upper"]
    assert g["left_bound"].tolist() == REF["left"] and g["right_bound"].tolist() == REF["right"]
    assert g["n_subcells"].tolist() == REF["nsub"] and g["cellArea1"].tolist() == REF["area"]
    assert g["cellCoor"].tolist() == REF["coor"]
    ids = g["lowres_id_on_highres"]  # [j][i]
    assert (ids[0:5, 0:5] == 1).all() and (ids[5:7, 0:5] == 3).all()
    assert (ids[0:5, 5:6] == 2).all() and (ids[5:7, 5:6] == 4).all()


def test_init_lowres_level_kat_oracle():
    L = orc_mpr._lib()
    m = grid_kat_mask()
    n0 = int(m.sum())
    nr, nc = C.c_int32(), C.c_int32()
    xll, yll, cs = C.c_double(), C.c_double(), C.c_double()
    L.orc_calculate_grid_properties(6, 7, 3973.0, 2735.0, 10.0, 50.0, C.byref(nr), C.byref(nc), C.byref(xll),
                                    C.byref(yll), C.byref(cs))
    assert (nr.value, nc.value, xll.value, yll.value, cs.value) == (2, 2, 3943.0, 2695.0, 50.0)
    mask1 = np.zeros(4, dtype=np.int32)
    coor = np.zeros(8, dtype=np.int32)
    area = np.full(n0, 25.0)
    a1 = np.zeros(4)
    up, lo, le, ri, ns = (np.zeros(4, dtype=np.int32) for _ in range(5))
    ids = np.zeros(42, dtype=np.int32)
    n1 = L.orc_init_lowres_level(6, 7, orc.iptr(m), orc.dptr(area), 10.0, 50.0, 2, 2, orc.iptr(mask1),
                                 orc.iptr(coor), orc.dptr(a1), orc.iptr(up), orc.iptr(lo), orc.iptr(le),
                                 orc.iptr(ri), orc.iptr(ns), orc.iptr(ids))
    assert n1 == 4
    assert lo.tolist() == REF["lower"] and up.tolist() == REF["upper"] and le.tolist() == REF["left"]
    assert ri.tolist() == REF["right"] and ns.tolist() == REF["nsub"] and a1.tolist() == REF["area"]
    assert coor.reshape(2, 4).tolist() == REF["coor"]


def test_library_grid_helper_equals_oracle_on_random_masks():
    L = orc_mpr._lib()
    rng = np.random.default_rng(5)
    for trial in range(30):
        nr0, nc0 = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        f = int(rng.integers(1, 7))
        m = (rng.random((nc0, nr0)) < rng.uniform(0.2, 0.95)).astype(np.int32)
        if m.sum() == 0:
            continue
        area = rng.uniform(1.0, 2.0, int(m.sum()))
        g = synth_mpr.init_lowres_level(m, 1.0, float(f), area)
        n1 = g["nCells1"]
        mask1 = np.zeros(g["nrows1"] * g["ncols1"], dtype=np.int32)
        coor = np.zeros(2 * n1, dtype=np.int32)
        a1 = np.zeros(n1)
        up, lo, le, ri, ns = (np.zeros(n1, dtype=np.int32) for _ in range(5))
        ids = np.zeros(nr0 * nc0, dtype=np.int32)
        n1o = L.orc_init_lowres_level(nr0, nc0, orc.iptr(m), orc.dptr(area), 1.0, float(f), g["nrows1"],
                                      g["ncols1"], orc.iptr(mask1), orc.iptr(coor), orc.dptr(a1), orc.iptr(up),
                                      orc.iptr(lo), orc.iptr(le), orc.iptr(ri), orc.iptr(ns), orc.iptr(ids))
        assert n1o == n1
        assert np.array_equal(up, g["upper_bound"]) and np.array_equal(lo, g["lower_bound"])
        assert np.array_equal(le, g["left_bound"]) and np.array_equal(ri, g["right_bound"])
        assert np.array_equal(ns, g["n_subcells"]) and np.array_equal(a1, g["cellArea1"])
        assert np.array_equal(ids, g["lowres_id_on_highres"].ravel())
        assert np.array_equal(mask1, g["mask1"].ravel())


def test_upscaling_operators_hand_checked():
    m = grid_kat_mask()
    n0 = int(m.sum())
    prob = {"nrows0": 6, "ncols0": 7, "mask0": m, "nL1": 4, "grid": synth_mpr.init_lowres_level(m, 10.0, 50.0)}
    # packed order = Fortran element order of the mask (numpy [j][i] raveled)
    x2d = np.arange(42, dtype=np.float64).reshape(7, 6) + 1.0
    x = x2d[m.astype(bool)]
    assert len(x) == n0
    ar = orc_mpr.upscale(prob, "arith", x)
    hm = orc_mpr.upscale(prob, "harm", x)
    gm = orc_mpr.upscale(prob, "geom", x)
    for k, (jl, jr, iu, idn) in enumerate(((0, 5, 0, 5), (0, 5, 5, 6), (5, 7, 0, 5), (5, 7, 5, 6))):
        sub = x2d[jl:jr, iu:idn][m[jl:jr, iu:idn].astype(bool)]
        assert np.isclose(ar[k], sub.mean(), rtol=1e-14)
        assert np.isclose(hm[k], len(sub) / (1.0 / sub).sum(), rtol=1e-14)
        assert np.isclose(gm[k], np.exp(np.log(sub).mean()), rtol=1e-12)
    cls = (np.arange(n0) % 3 + 1).astype(np.int32)
    fr = [orc_mpr.upscale(prob, "frac", cls, class_id=c) for c in (1, 2, 3)]
    assert np.allclose(sum(fr), 1.0)


@pytest.mark.parametrize("soil_case,pet_case", [(1, -1), (2, 0), (3, 1), (4, 2), (1, 3)])
def test_oracle_mpr_is_sane(soil_case, pet_case):
    prob = synth_mpr.make_mpr_problem(nx0=40, ny0=30, factor=5, nH=3, soil_case=soil_case, pet_case=pet_case)
    out = orc_mpr.run_mpr(prob)
    for k, v in out.items():
        assert np.isfinite(v).all(), k
    assert (out["L1_soilMoistFC"] <= out["L1_soilMoistSat"]).all()
    assert (out["L1_wiltingPoint"] <= out["L1_soilMoistFC"]).all()
    assert np.allclose(out["L1_fRoots"].sum(axis=1), 1.0)
    assert (out["L1_kFastFlow"] >= 1.0).all() and (out["L1_kSlowFlow"] >= out["L1_kFastFlow"]).all()
    assert (out["L1_kBaseFlow"] >= out["L1_kSlowFlow"]).all() and (out["L1_kPerco"] >= 2.0).all()
    assert ((out["L1_fSealed"] >= 0) & (out["L1_fSealed"] <= 0.6)).all()
    assert (out["L1_soilMoistSat"] > 1.0).all() and (out["L1_soilMoistSat"] < 1000.0).all()
    assert (out["L1_maxInter"] > 0).all()
    # a different land-cover scene changes scene-dependent fields only
    assert not np.array_equal(out["L1_fSealed"][0], out["L1_fSealed"][1])
