#!/usr/bin/env python
"""Generate tests/golden/test_domain_l0.npz: the L0 inputs of MPR for the reference's bundled
test basin exactly as the Fortran driver holds them after read_data / mpr_initialize, from the
ASCII grids and look-up tables under /root/reference/test_domain/input (container only).

The matching expected values are the L1 effective parameters the reference's own MPR wrote
into check/case_*/output_save/*_mHM_restart_*.nc, already stored in case_*.npz (param/*) by
make_golden.py together with each case's gamma vector (mhm_parameter.nml).

Restated here (input preparation only, not part of the product):
  * read_spatial_data_ascii transposition (common/mo_read_spatial_data.f90:222-227): the
    Fortran array is (file ncols, file nrows) with x fastest == numpy C order of loadtxt;
  * slope/aspect minimum values (MPR/mo_read_wrapper.f90:116-118,335-336);
  * empirical slope distribution (MPR/mo_mpr_startup.f90:278-310);
  * LAI per class from the LUT, clamped to [1e-10, 30] (MPR/mo_read_wrapper.f90:270-292);
  * soil data base for iFlag_soilDB = 0 incl. the tillage-depth horizon split and the depth
    weights Wd (MPR/mo_soil_database.f90:60-250, 387-500).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import h5lite  # noqa: E402

REF = "/root/reference"
INP = os.path.join(REF, "test_domain", "input")


def asc(path, dtype=np.float64):
    hdr = {}
    with open(path) as f:
        for _ in range(6):
            k, v = f.readline().split()
            hdr[k.lower()] = float(v)
    a = np.loadtxt(path, skiprows=6, dtype=dtype)
    assert a.shape == (int(hdr["nrows"]), int(hdr["ncols"]))
    return a, hdr


def soil_database(path, horizon_depth, tillage_depth):
    rows = []
    with open(path) as f:
        n_soil = int(f.readline().split()[1])
        f.readline()
        for line in f:
            p = line.split()
            if len(p) >= 7:
                rows.append((int(p[0]), int(p[1]), float(p[2]), float(p[3]), float(p[4]), float(p[5]), float(p[6])))
    nHor = np.zeros(n_soil, dtype=np.int32)
    nTill = np.full(n_soil, -9999, dtype=np.int32)
    RZ = np.full(n_soil, -9999.0)
    for ii, jj, up, down, cly, snd, bd in rows:
        nHor[ii - 1] = jj
        RZ[ii - 1] = max(RZ[ii - 1], np.rint(down))
    for ii, nH, up, down, cly, snd, bd in rows:
        if np.rint(up) < tillage_depth < np.rint(down):
            nHor[ii - 1] += 1
        if np.rint(up) < tillage_depth <= np.rint(down):
            nTill[ii - 1] = nH
    mh = int(nHor.max())
    UD = np.full((mh, n_soil), -9999.0)
    LD = np.full((mh, n_soil), -9999.0)
    clay = np.full((mh, n_soil), -9999.0)
    sand = np.full((mh, n_soil), -9999.0)
    dbm = np.full((mh, n_soil), -9999.0)
    kk = 0
    for jj, nH, up, down, cly, snd, bd in rows:
        s = jj - 1
        cly, snd = max(cly, 1.0), max(snd, 1.0)
        if cly + snd > 100.0:
            cly = cly / (cly + snd)
            snd = snd / (cly + snd)        # sic: uses the already rescaled clay (:214-217)
        if np.rint(up) < tillage_depth < np.rint(down):
            UD[nH - 1, s], LD[nH - 1, s] = np.rint(up), tillage_depth
            clay[nH - 1, s], sand[nH - 1, s], dbm[nH - 1, s] = cly, snd, bd
            up = tillage_depth
            kk = 1
        if kk == 1:
            nH += 1
        UD[nH - 1, s], LD[nH - 1, s] = np.rint(up), np.rint(down)
        clay[nH - 1, s], sand[nH - 1, s], dbm[nH - 1, s] = cly, snd, bd
        assert nH <= nHor[s]
        if nH == nHor[s]:
            kk = 0
    nHm = len(horizon_depth)
    hd = np.array(horizon_depth, dtype=np.float64)
    Wd = np.zeros((mh, nHm, n_soil))
    acc = 0.5
    for s in range(n_soil):
        Wd[nHor[s]:, :, s] = -9999.0
        hd[nHm - 1] = RZ[s]
        for jj in range(nHm):
            f = 0.0 if jj == 0 else hd[jj - 1]
            t = hd[jj] - acc
            lf = lt = -1
            for k in range(nHor[s]):
                if UD[k, s] <= f <= LD[k, s] - acc:
                    lf = k
                if UD[k, s] <= t <= LD[k, s] - acc:
                    lt = k
            assert 0 <= lf <= lt, (s, jj, lf, lt)
            if lf == lt:
                Wd[lf, jj, s] = 1.0
            else:
                Wd[lf, jj, s] = LD[lf, s] - f
                Wd[lt, jj, s] = (t + acc) - UD[lt, s]
                for k in range(lf + 1, lt):
                    Wd[k, jj, s] = LD[k, s] - UD[k, s]
                div = hd[jj] if jj == 0 else hd[jj] - hd[jj - 1]
                Wd[: nHor[s], jj, s] = Wd[: nHor[s], jj, s] / div
    return {"nSoil": n_soil, "maxHor": mh, "nHorizons": nHor, "nTillHorizons": nTill, "RZdepth": RZ,
            "sand": sand, "clay": clay, "DbM": dbm, "Wd": Wd, "HorizonDepth_last": hd[nHm - 1]}


def main():
    M = os.path.join(INP, "morph")
    dem, hdr = asc(os.path.join(M, "dem.asc"))
    mask0 = dem != hdr["nodata_value"]
    rst = h5lite.H5File(os.path.join(REF, "check", "case_00", "output_save", "b1_mHM_restart_001.nc"))
    assert np.array_equal(mask0, rst["L0_domain_mask"].read() != 0), "DEM mask == restart L0 mask"
    pick = lambda a: np.ascontiguousarray(a[mask0])
    n0 = int(mask0.sum())
    slope = np.maximum(pick(asc(os.path.join(M, "slope.asc"))[0]), 0.01)
    aspect = np.maximum(pick(asc(os.path.join(M, "aspect.asc"))[0]), 1.00)
    # empirical distribution of slope
    order = np.argsort(slope, kind="stable")
    emp = np.zeros(n0)
    emp[order[n0 - 1]] = n0 / (n0 + 1.0)
    for i in range(n0 - 2, -1, -1):
        a, b = order[i], order[i + 1]
        emp[a] = emp[b] if slope[a] == slope[b] else (i + 1) / (n0 + 1.0)
    soil = pick(asc(os.path.join(M, "soil_class.asc"), np.int64)[0]).astype(np.int32)
    geo = pick(asc(os.path.join(M, "geology_class.asc"), np.int64)[0]).astype(np.int32)
    lai_cls = pick(asc(os.path.join(M, "LAI_class.asc"), np.int64)[0]).astype(np.int32)
    lc = np.stack([pick(asc(os.path.join(INP, "luse", f), np.int64)[0]).astype(np.int32)
                   for f in ("lc_1981.asc", "lc_1991.asc")])
    # LAI look-up table
    ids, lut = [], []
    with open(os.path.join(M, "LAI_classdefinition.txt")) as f:
        n_lai_cls = int(f.readline().split()[1])
        f.readline()
        for _ in range(n_lai_cls):
            p = f.readline().split()
            ids.append(int(p[0]))
            lut.append([float(x) for x in p[2:14]])
    lut = np.array(lut)
    LAI0 = np.zeros((12, n0))
    for cid, row in zip(ids, lut):
        LAI0[:, lai_cls == cid] = row[:, None]
    LAI0 = np.clip(LAI0, 1.0e-10, 30.0)
    # geology look-up table
    gl, gk = [], []
    with open(os.path.join(M, "geology_classdefinition.txt")) as f:
        n_geo = int(f.readline().split()[1])
        f.readline()
        for _ in range(n_geo):
            p = f.readline().split()
            gl.append(int(p[1]))
            gk.append(int(p[2]))
    db = soil_database(os.path.join(M, "soil_classdefinition.txt"), [200.0, 0.0], 200.0)
    present = np.zeros(db["nSoil"], dtype=np.int32)
    present[np.unique(soil) - 1] = 1
    # river-network inputs (mRM/mo_mrm_read_data.f90:139-203): flow accumulation, flow direction
    # in the reference's rotated in-memory convention (:527-600), gauge locations, elevation
    rot = {1: 4, 2: 2, 4: 1, 8: 128, 16: 64, 32: 32, 64: 16, 128: 8}
    fdir_file = pick(asc(os.path.join(M, "fdir.asc"), np.int64)[0])
    fdir = np.array([rot.get(int(v), int(v)) for v in fdir_file], dtype=np.int32)
    facc = pick(asc(os.path.join(M, "facc.asc"), np.int64)[0]).astype(np.int32)
    gauges = pick(asc(os.path.join(M, "idgauges.asc"), np.int64)[0]).astype(np.int32)
    out = {"fDir0": fdir, "fAcc0": facc, "gaugeLoc0": gauges, "elev0": pick(dem),
           "xllcorner0": hdr["xllcorner"], "yllcorner0": hdr["yllcorner"],
           "mask0": mask0, "cellsize0": hdr["cellsize"], "geoUnit0": geo, "soilId0": soil, "LCover0": lc,
           "Asp0": aspect, "slope0": slope, "slope_emp0": emp, "y0": pick(rst["L0_domain_lat"].read()), "LAI0": LAI0,
           "GeoUnitList": np.array(gl, dtype=np.int32), "GeoUnitKar": np.array(gk, dtype=np.int32),
           "is_present": present, "fracSealed_CityArea": 0.6, "tillageDepth": 200.0}
    for k, v in db.items():
        out["soil/" + k] = v
    path = os.path.join(HERE, "test_domain_l0.npz")
    np.savez_compressed(path, **out)
    print("test_domain_l0: %d of %d x %d L0 cells, %d soil types (max %d horizons), %.0f kB" % (
        n0, mask0.shape[0], mask0.shape[1], db["nSoil"], db["maxHor"], os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
