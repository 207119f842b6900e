"""Forcing ingest: level-2 (meteo grid) chunk -> packed L1 forcing.
CPU: the oracle's restatement of spatial_aggregation / spatial_disaggregation against the
reference's own known-answer tests (src/tests/test_meteo_spatial_tools.pf:19-59, 117-159).
GPU: mhm_cuda_set_meteo_l2 (float64 and float32 input) against the oracle on random masks, and
a golden check case run from the 24 km meteo grid under a 12 km L1 grid."""
import ctypes as C

import numpy as np
import pytest

import golden_case
import orc
import parity

VAL_MASK = np.array([0, 1, 1, 0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1], dtype=np.int32)
NODATA = -9999.0


def l2_to_l1(data2, mask2, cs2, mask1, cs1):
    """numpy (nT, nc2, nr2) == Fortran (nr2, nc2, nT); returns (packed (nT, nCells1), grid (nT, nc1, nr1))"""
    L = orc.lib()
    pd, pi, i, d = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32, C.c_double
    L.orc_meteo_l2_to_l1.argtypes = [pd, i, i, i, pi, d, i, i, pi, d, pd, pd]
    L.orc_meteo_l2_to_l1.restype = i
    a = np.ascontiguousarray(data2, dtype=np.float64)
    m2, m1 = np.ascontiguousarray(mask2, dtype=np.int32), np.ascontiguousarray(mask1, dtype=np.int32)
    nT = a.shape[0]
    packed = np.zeros((nT, int(m1.sum())))
    grid = np.zeros((nT,) + m1.shape)
    n = L.orc_meteo_l2_to_l1(orc.dptr(a), m2.shape[1], m2.shape[0], nT, orc.iptr(m2), cs2, m1.shape[1], m1.shape[0],
                             orc.iptr(m1), cs1, orc.dptr(packed), orc.dptr(grid))
    assert n == packed.shape[1]
    return packed, grid


def test_spatial_aggregation_3d_kat():
    v = np.array([[0.21, 0.27, 0.21, 0.27, 0.26, 0.28, 0.28, 0.29, 0.32, 0.3],
                  [0.33, 0.32, 0.33, 0.32, 0.32, 0.33, 0.32, 0.34, 0.27, 0.23]])
    # Fortran data2(i, j, t) with mask2(i, j) = val_mask(j): numpy [t][j][i]
    data2 = np.zeros((2, 20, 20))
    for t in range(2):
        data2[t, :10, :] = v[t][:, None]
        data2[t, 10:, :] = v[t][:, None]
    mask2 = np.repeat(VAL_MASK[:, None], 20, axis=1)
    mask1 = np.array([[0, 1], [0, 1]], dtype=np.int32)       # reshape([F, T, F, T], [2, 2]) as [j][i]
    _, grid = l2_to_l1(data2, mask2, 1.0, mask1, 10.0)
    ref = np.full((2, 2, 2), NODATA)                          # [t][j][i]
    ref[0, 0, 1], ref[1, 0, 1], ref[0, 1, 1], ref[1, 1, 1] = 0.26, 0.293, 0.275, 0.306
    assert np.allclose(grid, ref, atol=1e-3, rtol=0)


def test_spatial_disaggregation_3d_kat():
    data2 = np.zeros((2, 2, 2))                               # [t][j][i]
    data2[0, 0, 0], data2[1, 0, 0], data2[0, 0, 1], data2[1, 0, 1] = 0.21, 0.27, 0.26, 0.28
    data2[0, 1, 0], data2[0, 1, 1], data2[1, 1, 0], data2[1, 1, 1] = 0.21, 0.28, 0.27, 0.29
    mask1 = np.repeat(VAL_MASK[:, None], 20, axis=1)          # mask1(i, j) = val_mask(j)
    mask2 = np.array([[0, 1], [0, 1]], dtype=np.int32)
    _, grid = l2_to_l1(data2, mask2, 10.0, mask1, 1.0)
    ref = np.full((2, 20, 20), NODATA)
    ref[0, :10, 10:], ref[1, :10, 10:] = 0.26, 0.28           # i = 11..20, j = 1..10
    ref[0, 10:, 10:], ref[1, 10:, 10:] = 0.28, 0.29
    assert np.allclose(grid, ref, atol=1e-3, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["aggregate", "disaggregate", "equal"])
def test_device_ingest_equals_oracle(mode):
    from mhm_b200 import interface, synth

    rng = np.random.default_rng(4)
    if mode == "aggregate":
        nr1, nc1, f = 13, 9, 4
        nr2, nc2, cs1, cs2 = nr1 * f - 2, nc1 * f - 1, 4000.0, 1000.0   # ragged last rows / columns
    elif mode == "disaggregate":
        nr2, nc2, f = 7, 5, 3
        nr1, nc1, cs1, cs2 = nr2 * f - 1, nc2 * f, 1000.0, 3000.0
    else:
        nr1 = nr2 = 17
        nc1 = nc2 = 11
        cs1 = cs2 = 2000.0
    mask2 = (rng.random((nc2, nr2)) < 0.9).astype(np.int32)
    if mode == "aggregate":      # an L1 cell is valid where it holds at least one valid level-2 cell
        mask1 = np.zeros((nc1, nr1), dtype=np.int32)
        js, is_ = np.nonzero(mask2)
        mask1[js // f, is_ // f] = 1
    elif mode == "disaggregate":
        mask1 = np.repeat(np.repeat(mask2, f, axis=0), f, axis=1)[:nc1, :nr1].copy()
    else:
        mask1 = mask2.copy()
    nT, n1 = 6, int(mask1.sum())
    data2 = rng.uniform(-5.0, 30.0, (nT, nc2, nr2))
    want, _ = l2_to_l1(data2, mask2, cs2, mask1, cs1)
    pm = synth.process_matrix(1, -1, 0)
    with interface.Context() as ctx:
        dom = ctx.register_domain(1, n1, 2, 12, 2, pm)
        dom.set_meteo_config(-1, 24, True, False, synth.FNIGHT_PREC, synth.FNIGHT_PET, synth.FNIGHT_TEMP,
                             synth.EVAP_COEFF)
        dom.set_meteo_l2("pre", data2, mask2, cs2, mask1, cs1)
        got = dom.get_meteo("pre", nT)
        parity.assert_bit_exact(got, want, "float64 chunk, " + mode)
        d32 = data2.astype(np.float32)
        want32, _ = l2_to_l1(d32.astype(np.float64), mask2, cs2, mask1, cs1)
        dom.set_meteo_l2("temp", d32, mask2, cs2, mask1, cs1)
        parity.assert_bit_exact(dom.get_meteo("temp", nT), want32, "float32 chunk, " + mode)


@pytest.mark.gpu
def test_reference_run_from_the_meteo_grid():
    """check/case_04 domain 5 (L1 = 12 km) fed with the 24 km meteo grid as it is in the files:
    the disaggregation happens on the device; results against the reference run"""
    import orc_run  # noqa: F401  (oracle build)
    from mhm_b200 import driver, interface

    prob, ref = golden_case.load("case_04_b5")
    z = np.load(golden_case.HERE + "/golden/case_04_b5.npz")
    mask1 = z["mask1"]
    # rebuild the level-2 chunks: every 2 x 2 block of L1 cells holds its parent's value
    mask2 = np.zeros((mask1.shape[0] // 2, mask1.shape[1] // 2), dtype=np.int32)
    js, is_ = np.nonzero(mask1)
    mask2[js // 2, is_ // 2] = 1
    with interface.Context() as ctx:
        ctx.set_math_mode("strict")
        dom = driver.setup_domain(ctx, 1, prob, upload_forcing=False)
        for var in ("pre", "temp", "pet"):
            F = prob["forcing"][var]
            l2 = np.zeros((F.shape[0],) + mask2.shape, dtype=np.float32 if var != "pre" else np.float64)
            l2[:, js // 2, is_ // 2] = F
            if l2.dtype == np.float32:
                assert np.array_equal(l2[:, js // 2, is_ // 2].astype(np.float64), F), "file values are float32-exact"
            dom.set_meteo_l2(var, l2, mask2, 24000.0, mask1, 12000.0)
        dom.run_steps(1, prob["time"]["nTimeSteps"])
        q = golden_case.daily_mean(dom.get_runoff(), ref["warming_days"])
        parity.assert_close(q, ref["Qsim"], "daily discharge", rtol=parity.RTOL_Q)
        parity.assert_close(dom.get_variable("L1_soilMoist"), ref["final"]["L1_soilMoist"], "soil moisture")
