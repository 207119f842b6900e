"""River-network initialisation (mhm_b200/csrc/netinit.cu, host only) against the
reference's own mRM restart files: flow direction at L11, draining cells, link topology, routing
order, link locations on the L0 grid, gauge nodes, link length, slope, flood-plain area and its
impervious fraction -- all bit-identical,
for the bundled test basin at 24 km and 12 km routing resolution."""
import numpy as np
import pytest

import golden_case
from mhm_b200 import netinit, synth_mpr

PAIRS = [("fDir11", "L11_fDir"), ("rowOut", "L11_rowOut"), ("colOut", "L11_colOut"), ("fromN", "L11_fromN"),
         ("toN", "L11_toN"), ("rOrder", "L11_rOrder"), ("netPerm", "L11_netPerm"), ("fRow", "L11_fRow"),
         ("fCol", "L11_fCol"), ("tRow", "L11_tRow"), ("tCol", "L11_tCol")]


@pytest.mark.parametrize("case,res11", [("case_00", 24000.0), ("case_10", 24000.0), ("case_04_b1", 24000.0),
                                        ("case_04_b2", 24000.0), ("case_04_b4", 24000.0), ("case_04_b5", 12000.0)])
def test_network_equals_reference_restart(case, res11):
    z0 = np.load(golden_case.HERE + "/golden/test_domain_l0.npz")
    zc = np.load(golden_case.HERE + "/golden/%s.npz" % case)
    n0 = int(z0["mask0"].sum())
    g = synth_mpr.init_lowres_level(z0["mask0"], float(z0["cellsize0"]), res11, np.full(n0, float(z0["cellsize0"]) ** 2))
    assert np.array_equal(g["mask1"] != 0, zc["net/mask11"])
    r = netinit.net_init(z0["mask0"], z0["fDir0"], z0["fAcc0"], z0["elev0"], float(z0["cellsize0"]), g,
                         z0["gaugeLoc0"], [398], xll=float(z0["xllcorner0"]), yll=float(z0["yllcorner0"]),
                         LCover0=z0["LCover0"])
    nl = r["nLinks"]
    assert nl == int((zc["net/L11_fromN"] > 0).sum()) and r["nOutlets11"] == g["nCells1"] - nl
    for ours, theirs in PAIRS:
        n = g["nCells1"] if ours in ("fDir11", "rowOut", "colOut") else nl
        assert np.array_equal(r[ours][:n], zc["net/" + theirs][:n]), ours
    assert np.array_equal(r["gaugeNodeList"], zc["net/gaugeNodeList"])
    # link length [m] and slope: same operations in the same order -> identical doubles
    assert np.array_equal(r["length"][:nl], zc["net/L11_length"][:nl])
    assert np.array_equal(r["slope"][:nl], zc["net/L11_slope"][:nl])
    # flood plains: area per link and its impervious share per land-cover scene (the reference
    # evaluates the latter for nLinks + 1 entries, the last one is -0.0 / nodata)
    assert np.array_equal(r["aFloodPlain"][:nl], zc["net/L11_aFloodPlain"][:nl])
    got, want = r["nLinkFracFPimp"][:, : nl + 1], zc["net/L11_nLinkFracFPimp"][:, : nl + 1]
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # every L0 cell drains to the node whose draining cell it reaches first
    assert (r["draCell0"] >= 1).all() and (r["draCell0"] <= g["nCells1"]).all()
    assert r["L0_nOutlets"] == 1 and (r["draSC0"] > 0).sum() == g["nCells1"]


def test_network_init_scales_linearly_and_feeds_the_routing():
    """a 400 x 300 synthetic D8 grid (tilted plane + noise, pits filled by construction): the
    network comes out as a forest whose netPerm is a valid topological order"""
    rng = np.random.default_rng(11)
    ny, nx, f = 300, 400, 4   # numpy (ncols0, nrows0): x is the fast index
    mask0 = np.ones((ny, nx), dtype=bool)
    # every cell drains east / south-east / north-east: in-memory codes 4 (i+1), 2 (i+1, j+1), 128 (i-1, j+1)?
    # use codes that move the FIRST index up: 4 = (i+1, j), 2 = (i+1, j+1), 8 = (i+1, j-1)
    codes = rng.choice(np.array([4, 2, 8], dtype=np.int32), size=(ny, nx))
    codes[0, codes[0] == 8] = 4      # stay inside at j = 1
    codes[-1, codes[-1] == 2] = 4    # ... and at j = ncols0
    # flow accumulation by a sweep in x
    acc = np.ones((ny, nx), dtype=np.int64)
    for i in range(nx - 1):
        for dj, c in ((0, 4), (1, 2), (-1, 8)):
            src = np.nonzero(codes[:, i] == c)[0]
            np.add.at(acc[:, i + 1], src + dj, acc[src, i])
    elev = (nx - np.arange(nx))[None, :] * 1.0 + rng.random((ny, nx)) * 0.1
    n0 = ny * nx
    g = synth_mpr.init_lowres_level(mask0, 100.0, 100.0 * f, np.full(n0, 1.0e4))
    r = netinit.net_init(mask0, codes.ravel(), acc.ravel().astype(np.int32), elev.ravel(), 100.0, g)
    nn, nl = g["nCells1"], r["nLinks"]
    assert r["L0_nOutlets"] == ny and r["nOutlets11"] >= 1 and nl == nn - r["nOutlets11"]
    seen = np.zeros(nn + 1, dtype=bool)
    has_up = np.zeros(nn + 1, dtype=bool)
    has_up[r["toN"][:nl]] = True
    pos = np.zeros(nl, dtype=np.int64)
    pos[r["netPerm"][:nl] - 1] = np.arange(nl)
    link_of = np.full(nn + 1, -1)
    link_of[r["fromN"][:nl]] = np.arange(nl)
    down = link_of[r["toN"][:nl]]
    ok = down < 0
    assert (pos[np.where(ok, 0, down)][~ok] > pos[~ok]).all()   # a link is routed before the link below it
    assert (r["length"][:nl] > 0).all() and (r["slope"][:nl] >= 0.0001).all()


def _test_basin(res11=24000.0, routingCase=1):
    z0 = np.load(golden_case.HERE + "/golden/test_domain_l0.npz")
    n0 = int(z0["mask0"].sum())
    g = synth_mpr.init_lowres_level(z0["mask0"], float(z0["cellsize0"]), res11, np.full(n0, float(z0["cellsize0"]) ** 2))
    r = netinit.net_init(z0["mask0"], z0["fDir0"], z0["fAcc0"], z0["elev0"], float(z0["cellsize0"]), g,
                         z0["gaugeLoc0"], [398], xll=float(z0["xllcorner0"]), yll=float(z0["yllcorner0"]),
                         LCover0=z0["LCover0"], routingCase=routingCase)
    return z0, g, r


def test_flow_accumulation_equals_reference_restart():
    """L11_flow_accumulation against L11_fAcc of the reference's mRM restart files (24 and 12 km)"""
    for case, res11 in (("case_00", 24000.0), ("case_04_b5", 12000.0)):
        z0, g, r = _test_basin(res11)
        zc = np.load(golden_case.HERE + "/golden/%s.npz" % case)
        facc = netinit.flow_accumulation(g, r["fDir11"])
        assert np.array_equal(facc, zc["net/L11_fAcc"]), case


def test_celerity_routing_parameters_equal_reference_and_oracle():
    """routing cases 2 and 3 from the raw grids: link-length floor, mrm_update_param and
    L11_calc_celerity.  Case 2 against the reference's restart file of check/case_09 (length, C1,
    C2, TSrout bit-identical); case 3 against the oracle's literal restatement (which
    check/case_13's discharge pins, tests/test_golden_reference.py)."""
    import orc_run

    z0, g, r = _test_basin(routingCase=2)
    z9 = np.load(golden_case.HERE + "/golden/case_09.npz")
    nn, nl = g["nCells1"], r["nLinks"]
    assert np.array_equal(r["length"], z9["net/L11_length"])          # floored at the 40th percentile
    c1, c2, ts = netinit.update_param(r["length"], float(z9["celerity"][0]), nn - nl)
    assert ts == float(z9["net/L11_TSrout"][0])
    assert np.array_equal(c1[:nl], z9["final/L11_C1"][:nl]) and np.array_equal(c2[:nl], z9["final/L11_C2"][:nl])

    z0, g, r3 = _test_basin(routingCase=3)
    prob, _ = golden_case.load("case_13")
    onet = orc_run.case3_params(prob["net"], z0)                      # oracle: from the reference's link locations
    assert np.array_equal(r3["streamNet0"], onet["streamNet0"])
    assert np.array_equal(r3["length"], onet["L11_length"])
    cel11, cel0 = netinit.calc_celerity(z0["mask0"], z0["fDir0"], z0["slope0"], r3, onet["slope_factor"])
    assert np.array_equal(cel11[:nl], onet["L11_celerity"][:nl])
    on_stream = r3["streamNet0"] > 0
    assert (cel0[on_stream] > 0).all() and (cel0[~on_stream] == -9999.0).all()
    c1, c2, ts = netinit.update_param(r3["length"], np.where(cel11 > 0, cel11, -9999.0), nn - nl)
    assert ts == onet["TSrout"] == 7200.0
    assert np.array_equal(c1[:nl], onet["C1"][:nl]) and np.array_equal(c2[:nl], onet["C2"][:nl])


def test_flow_accumulation_linear_on_a_deep_network():
    """a 600 x 500 synthetic grid (chains thousands of cells deep): the explicit post-order walk
    equals a topological-order sum and needs no recursion"""
    rng = np.random.default_rng(3)
    ny, nx = 500, 600                              # numpy (ncols, nrows)
    codes = rng.choice(np.array([4, 2, 8], dtype=np.int32), size=(ny, nx))   # all move the first index up
    codes[0, codes[0] == 8] = 4
    codes[-1, codes[-1] == 2] = 4
    codes[:, -1] = 0                               # sinks on the last row of the first index
    mask = np.ones((ny, nx), dtype=np.int32)
    area = rng.integers(1, 5, ny * nx).astype(np.float64) * 1.0e6
    g = {"mask1": mask, "nrows1": nx, "ncols1": ny, "nCells1": ny * nx, "cellArea1": area}
    got = netinit.flow_accumulation(g, codes.ravel()).reshape(ny, nx)
    f = float(np.float32(1.e-6))
    acc = area.reshape(ny, nx) * f
    for i in range(nx - 1):                        # integers x f: every partial sum is exact enough to compare closely
        for dj, c in ((0, 4), (1, 2), (-1, 8)):
            src = np.nonzero(codes[:, i] == c)[0]
            np.add.at(acc[:, i + 1], src + dj, acc[src, i])
    assert np.allclose(got, acc, rtol=1e-12, atol=0.0)
