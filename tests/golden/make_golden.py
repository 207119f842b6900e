#!/usr/bin/env python
"""Generate the golden fixtures tests/golden/case_*.npz from the reference's OWN saved outputs.

Runs only in the build container (reads /root/reference; the GPU box has no such path) and is
committed together with the fixtures it produced.  For every check case it collects
  * the inputs of the hot path as the reference itself saw them: the daily forcing of the
    simulation period (test_domain/input/meteo/*.nc, packed with the L1 mask), the effective
    parameters the reference's MPR produced (output_save/*_mHM_restart_*.nc), the river network
    integers and link properties (output_save/*_mRM_restart_*.nc), gamma (mhm_parameter.nml);
  * the outputs of the reference run: final states and last-step fluxes (mHM restart), final
    routing state + C1/C2/K/xi (mRM restart), the daily gauge discharge in full double precision
    (output_save/*_discharge.nc) and its 7-decimal text twin (*_daily_discharge.out).
The .nc files are NetCDF-4/HDF5 and are read with the pure-Python reader h5lite.py.

    python tests/golden/make_golden.py            # writes tests/golden/case_00.npz ...
"""
import datetime
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import h5lite  # noqa: E402

REF = "/root/reference"

L1_PARAMS_LC = ["L1_fSealed", "L1_alpha", "L1_degDayInc", "L1_degDayMax", "L1_degDayNoPre",
                "L1_kfastFlow", "L1_kSlowFlow", "L1_kBaseFlow", "L1_kPerco", "L1_tempThresh"]
L1_PARAMS_LC_H = ["L1_fRoots", "L1_soilMoistFC", "L1_soilMoistSat", "L1_soilMoistExp",
                  "L1_wiltingPoint"]
L1_PARAMS_1 = ["L1_karstLoss", "L1_unsatThresh", "L1_sealedThresh", "L1_fAsp", "L1_HarSamCoeff",
               "L1_jarvis_thresh_c1"]
L1_PARAMS_LAI = ["L1_maxInter", "L1_PrieTayAlpha", "L1_surfResist"]
L1_PARAMS_LC_LAI = ["L1_petLAIcorFactor", "L1_aeroResist"]
L1_STATES = ["L1_Inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW"]
L1_FLUXES = ["L1_aETCanopy", "L1_aETSealed", "L1_baseflow", "L1_fastRunoff", "L1_melt",
             "L1_percol", "L1_preEffect", "L1_rain", "L1_runoffSeal", "L1_slowRunoff", "L1_snow",
             "L1_Throughfall", "L1_total_runoff", "L1_degDay"]
RENAME = {"L1_kfastFlow": "L1_kFastFlow", "L1_Inter": "L1_inter"}


def nml_values(path, group, key):
    """numbers of `key = ...` inside &group of a namelist file (continuation lines included)"""
    txt = open(path).read()
    m = re.search(r"&%s\b(.*?)^\s*/" % group, txt, re.S | re.M | re.I)
    body = re.sub(r"!.*", "", m.group(1))
    m = re.search(r"^\s*%s\s*=\s*(.*?)(?=^\s*[A-Za-z_][\w%%()\s:,]*=|\Z)" % re.escape(key), body,
                  re.S | re.M | re.I)
    return [float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eEdD][-+]?\d+)?", m.group(1).replace("d", "e"))]


def gamma_group(par_nml, group):
    """values of a group of mhm_parameter.nml: each line is lower, upper, value, flag, scaling"""
    txt = open(par_nml).read()
    m = re.search(r"&%s\b(.*?)^\s*/" % group, txt, re.S | re.M | re.I)
    vals = []
    for line in m.group(1).splitlines():
        line = re.sub(r"!.*", "", line)
        if "=" in line:
            nums = [float(x) for x in line.split("=")[1].replace(",", " ").split()]
            vals.append(nums[2])
    return np.array(vals)


def gamma_vector(par_nml, cases):
    """the flat global parameter vector and processMatrix (numpy (3, 11)) for a process selection,
    assembled in process order like MPR/mo_mpr_read_config.f90:390-988"""
    soil, pet, rout = cases
    groups = ["interception1", "snow1", "soilmoisture%d" % soil, "directRunoff1",
              {-1: "PETminus1", 0: "PET0", 1: "PET1", 2: "PET2", 3: "PET3"}[pet], "interflow1",
              "percolation1", {0: None, 1: "routing1", 2: "routing2", 3: "routing3"}[rout], "geoparameter"]
    case_of = [1, 1, soil, 1, pet, 1, 1, rout, 1]
    pm = np.zeros((3, 11), dtype=np.int32)
    vals, end = [], 0
    for p, (g, c) in enumerate(zip(groups, case_of)):
        v = [] if g is None else list(gamma_group(par_nml, g))
        vals += v
        end += len(v)
        pm[:, p] = [c, len(v), end]
    pm[:, 9] = [0, 0, end]
    pm[:, 10] = [0, 0, end]
    return np.array(vals), pm


def make_case(case, sim_start, n_days, warming_days, lc_years, cases=(1, 0, 1), out_prefix="b1",
              restart_id="001", name=None, net_from=None):
    """cases = processCase(3) soil moisture, (5) PET, (8) routing of the check case's mhm.nml.
    net_from: the check case saved no mRM restart; the (identical) river network comes from that
    case's restart, without its final routing state and Muskingum parameters."""
    cdir = os.path.join(REF, "check", case)
    sav = os.path.join(cdir, "output_save")
    mhm = h5lite.H5File(os.path.join(sav, "%s_mHM_restart_%s.nc" % (out_prefix, restart_id)))
    routing = cases[2] != 0
    mrm_dir = sav if net_from is None else os.path.join(REF, "check", net_from, "output_save")
    mrm = h5lite.H5File(os.path.join(mrm_dir, "%s_mRM_restart_%s.nc" % (out_prefix, restart_id))) if routing else None
    out = {"cases": np.array(cases, dtype=np.int32)}
    mask1 = mhm["L1_domain_mask"].read() != 0           # nc (y, x) == Fortran (x, y), x fastest
    # restart masks: 1 = valid?  check against the cell count below
    n = int(mask1.sum())
    pick = lambda a: np.ascontiguousarray(a[..., mask1])
    out["mask1"] = mask1
    P = {}
    for k in L1_PARAMS_LC:
        if k in mhm:
            P[RENAME.get(k, k)] = pick(mhm[k].read())[:, None, :]
    for k in L1_PARAMS_LC_H + L1_PARAMS_LC_LAI:
        if k in mhm:
            P[k] = pick(mhm[k].read())
    for k in L1_PARAMS_1:
        if k in mhm:
            P[k] = pick(mhm[k].read())[None, None, :]
    for k in L1_PARAMS_LAI:
        if k in mhm:
            P[k] = pick(mhm[k].read())[None, :, :]
    for k, v in P.items():
        assert np.all(v != -9999.0), k
        out["param/" + k] = v
    for k in L1_STATES + ["L1_soilMoist"] + L1_FLUXES + ["L1_aETSoil", "L1_infilSoil"]:
        out["final/" + RENAME.get(k, k)] = pick(mhm[k].read())
    out["horizon_bnds"] = mhm["L1_SoilHorizons_bnds"].read()
    out["L1_areaCell_km2"] = pick(mhm["L1_domain_cellarea"].read())
    out["L1_lat"] = pick(mhm["L1_domain_lat"].read())
    if routing:
        network(out, mrm, pick, cdir)
        if net_from is not None:
            for k in [k for k in out if k.startswith("final/L11_")] + ["net/L11_TSrout", "net/ProcessMatrix"]:
                del out[k]
            out["net_from"] = np.array(net_from)
        if cases[2] == 3:
            out["slope_factor"] = gamma_group(os.path.join(cdir, "mhm_parameter.nml"), "routing3")
    forcing(out, mask1, pick, sim_start, n_days, cases[1])
    if routing:
        q = h5lite.H5File(os.path.join(sav, "%s_discharge.nc" % out_prefix))
        for k in q.keys():
            if k.startswith("Qsim_"):
                out["Qsim/" + k[5:]] = q[k].read()
        if os.path.exists(os.path.join(sav, "%s_daily_discharge.out" % out_prefix)):
            txt = np.loadtxt(os.path.join(sav, "%s_daily_discharge.out" % out_prefix), skiprows=1)
            out["Qsim_text"] = txt[:, 5::2]
    # gridded outputs of the run (mhm_outputs.nml: outputFlxState, timeStep_model_outputs)
    fs = os.path.join(sav, "%s_mHM_Fluxes_States.nc" % out_prefix)
    if os.path.exists(fs):
        txt = re.sub(r"!.*", "", open(os.path.join(cdir, "mhm_outputs.nml")).read())
        flags = np.zeros(21, dtype=np.int32)
        for m in re.finditer(r"outputFlxState\((\d+)\)\s*=\s*\.(TRUE|FALSE)\.", txt, re.I):
            flags[int(m.group(1)) - 1] = m.group(2).upper() == "TRUE"
        out["out_flags"] = flags
        out["out_timestep"] = np.array(int(re.search(r"timeStep_model_outputs\s*=\s*(-?\d+)", txt, re.I).group(1)))
        f = h5lite.H5File(fs)
        for k in f.keys():
            if len(f[k].shape) == 3:
                out["out/" + k] = pick(f[k].read())       # (windows, nCells)
        out["out_time_bnds"] = f["time_bnds"].read()
    out["gamma"], out["processMatrix"] = gamma_vector(os.path.join(cdir, "mhm_parameter.nml"), cases)
    out["time"] = np.array([sim_start.toordinal(), n_days, warming_days])
    out["lc_years"] = np.array(lc_years, dtype=np.int32)   # (first year, LCyearId...)
    name = name or case
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%s: %d L1 cells, %s nodes, %d forcing days, %d keys, %.0f kB" % (
        name, n, int(out["net/mask11"].sum()) if routing else "no", n_days, len(out),
        os.path.getsize(path) / 1e3))


def network(out, mrm, pick, cdir):
    # ---- network ------------------------------------------------------------------------
    mask11 = mrm["L11_domain_mask"].read() != 0
    pick11 = lambda a: np.ascontiguousarray(a[..., mask11])
    for k in ("L11_fromN", "L11_toN", "L11_netPerm", "L11_rOrder", "L11_label", "L11_sink",
              "L11_length", "L11_slope", "L11_aFloodPlain", "gaugeNodeList", "L11_TSrout",
              "ProcessMatrix", "L11_fRow", "L11_fCol", "L11_tRow", "L11_tCol"):
        out["net/" + k] = mrm[k].read()
    for k in ("L11_Id", "L11_fDir", "L11_fAcc", "L11_rowOut", "L11_colOut", "L11_L1_Id",
              "L11_nLinkFracFPimp"):
        out["net/" + k] = pick11(mrm[k].read())
    out["net/L1_L11_Id"] = pick(mrm["L1_L11_Id"].read())
    out["net/L1_Id"] = pick(mrm["L1_Id"].read())
    out["net/L11_areaCell_km2"] = pick11(mrm["L11_domain_cellarea"].read())
    out["net/mask11"] = mask11
    for k in ("L11_Qmod", "L11_qOUT", "L11_qTIN", "L11_qTR", "L11_K", "L11_xi", "L11_C1", "L11_C2"):
        out["final/" + k] = pick11(mrm[k].read())
    out["rout_param"] = gamma_group(os.path.join(cdir, "mhm_parameter.nml"), "routing1")
    out["celerity"] = gamma_group(os.path.join(cdir, "mhm_parameter.nml"), "routing2")


def forcing(out, mask1, pick, sim_start, n_days, pet_case):
    """daily forcing of the simulation period on the L1 cells.  The test domain's meteo grid is
    24 km; a 12 km L1 grid takes the value of its parent cell (ic = ceiling(i / cellFactor),
    meteo/mo_meteo_spatial_tools.f90:361-369), an equal grid takes it as it is."""
    met = os.path.join(REF, "test_domain", "input", "meteo")
    files = [("pre", "pre/pre.nc", "pre"), ("tavg", "tavg/tavg.nc", "temp")]
    if pet_case in (-1, 0):
        files.append(("pet", "pet/pet.nc", "pet"))
    if pet_case == 1:
        files += [("tmin", "tmin.nc", "tmin"), ("tmax", "tmax.nc", "tmax")]
    if pet_case in (2, 3):
        files.append(("net_rad", "net_rad.nc", "netrad"))
    if pet_case == 3:
        files += [("eabs", "eabs.nc", "absvappress"), ("windspeed", "windspeed.nc", "windspeed")]
    for var, rel, key in files:
        f = h5lite.H5File(os.path.join(met, rel))
        t = f["time"].read()
        units = f["time"].attrs["units"]
        base = datetime.date(*[int(x) for x in re.search(r"since (\d+)-(\d+)-(\d+)", units).groups()])
        assert units.startswith("days since")
        i0 = int(np.nonzero(t == (sim_start - base).days)[0][0])
        a = f[var].read()[i0:i0 + n_days].astype(np.float64)
        assert a.shape[0] == n_days
        fy, fx = mask1.shape[0] // a.shape[1], mask1.shape[1] // a.shape[2]
        assert fy == fx and a.shape[1] * fy == mask1.shape[0] and a.shape[2] * fx == mask1.shape[1]
        if fy > 1:
            a = np.repeat(np.repeat(a, fy, axis=1), fx, axis=2)
        a = pick(a)
        assert np.all(a != -9999.0)
        out["forcing/" + key] = a


def make_case_optimised(case, sim_start, n_days, warming_days, lc_years, cases):
    """a calibration check case: no mHM restart is written, the saved discharge belongs to the
    final run with the optimiser's best parameter set (output_save/FinalParam.out, 15 digits).
    The fixture holds forcing, that gamma, the network / final routing state and Qsim; the L1
    parameters have to come from MPR."""
    cdir = os.path.join(REF, "check", case)
    sav = os.path.join(cdir, "output_save")
    mrm = h5lite.H5File(os.path.join(sav, "b1_mRM_restart_001.nc"))
    ref00 = h5lite.H5File(os.path.join(REF, "check", "case_00", "output_save", "b1_mHM_restart_001.nc"))
    mask1 = ref00["L1_domain_mask"].read() != 0
    pick = lambda a: np.ascontiguousarray(a[..., mask1])
    out = {"cases": np.array(cases, dtype=np.int32), "mask1": mask1,
           "L1_lat": pick(ref00["L1_domain_lat"].read()), "horizon_bnds": ref00["L1_SoilHorizons_bnds"].read()}
    network(out, mrm, pick, cdir)
    forcing(out, mask1, pick, sim_start, n_days, cases[1])
    q = h5lite.H5File(os.path.join(sav, "b1_discharge.nc"))
    for k in q.keys():
        if k.startswith("Qsim_"):
            out["Qsim/" + k[5:]] = q[k].read()
    txt = np.loadtxt(os.path.join(sav, "b1_daily_discharge.out"), skiprows=1)
    out["Qsim_text"] = txt[:, 5::2]
    # best parameter set: second line of FinalParam.out = objective followed by all parameters
    with open(os.path.join(sav, "FinalParam.out")) as f:
        names = f.readline().split()
        vals = [float(x) for x in f.readline().split()]
    assert names[0] == "OF" and len(names) == len(vals)
    _, pm = gamma_vector(os.path.join(cdir, "mhm_parameter.nml"), cases)
    assert pm[2, 8] == len(vals) - 1, (pm[2], len(vals))
    out["gamma"], out["processMatrix"] = np.array(vals[1:]), pm
    out["rout_param"] = out["gamma"][pm[2, 7] - pm[1, 7]: pm[2, 7]]
    out["time"] = np.array([sim_start.toordinal(), n_days, warming_days])
    out["lc_years"] = np.array(lc_years, dtype=np.int32)
    path = os.path.join(HERE, case + ".npz")
    np.savez_compressed(path, **out)
    print("%s: optimised run, %d parameters, %d forcing days, %.0f kB" % (case, len(vals) - 1, n_days,
                                                                        os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    D = datetime.date
    LC = lambda y0, ny: [y0] + [1 if y <= 1990 else 2 for y in range(y0, y0 + ny)]
    # check/case_00/mhm.nml: warming 181 d before 1990-07-01, evaluation until 1991-06-30;
    # land cover scenes 1981-1990 -> 1, 1991-2000 -> 2; PET as input with aspect correction,
    # Feddes, Muskingum routing case 1
    make_case("case_00", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (1, 0, 1))
    # Jarvis soil moisture + Priestley-Taylor PET, no routing
    make_case("case_02", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (2, 2, 0))
    # Hargreaves-Samani PET + routing case 2 (constant celerity, adaptive routing step)
    make_case("case_09", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (1, 1, 2))
    # soil moisture case 3 (Jarvis, FC-dependent roots) / 4 (Feddes, FC-dependent roots) + LAI-corrected PET
    make_case("case_10", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (3, -1, 1))
    make_case("case_12", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (4, -1, 1))
    # routing case 3 (celerity from the river slope) + river temperature (not part of the path);
    # the run saved no mRM restart: same domain and routing resolution as case_00
    make_case("case_13", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (1, 0, 3), net_from="case_00")
    # DDS calibration with Penman-Monteith PET (processCase(5) = 3): final run with the best set
    make_case_optimised("case_03", D(1990, 1, 1), 181 + 365, 181, LC(1990, 2), (1, 3, 1))
    # case_04: six domains; 1, 2, 4, 5 use the test domain (3 and 6 need forcing files that
    # are not in the tree).  2: L1 12 km under L11 24 km; 5: L1 = L11 = 12 km
    e = lambda a, b, w: (a - datetime.timedelta(days=w), (b - a).days + 1 + w, w)
    for dom, (a, b, w) in {1: (D(1990, 7, 1), D(1990, 12, 31), 180), 2: (D(1991, 1, 1), D(1992, 12, 31), 180),
                           4: (D(1990, 7, 1), D(1990, 12, 31), 360), 5: (D(1992, 1, 1), D(1992, 12, 31), 50)}.items():
        s0, nd, w = e(a, b, w)
        make_case("case_04", s0, nd, w, LC(s0.year, b.year - s0.year + 1), (1, 0, 1),
                  out_prefix="b%d" % dom, restart_id="%03d" % dom, name="case_04_b%d" % dom)
