"""Synthetic domains of the BASELINE shapes: parameters inside the post-MPR
physical ranges, hourly or daily forcing, and a Scheidegger-type river network.

Everything is numpy, seeded, and laid out like the reference's module globals:
Fortran (nCells, dim2, dim3) == numpy C-order (dim3, dim2, nCells).
"""
import numpy as np

SEED = 20261017
JUL_1990_01_01 = 2447893  # julday(1, 1, 1990)

# mhm.nml: &panEvapo, &nightDayRatio
EVAP_COEFF = [1.30, 1.20, 0.72, 0.75, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.00, 1.50]
FNIGHT_PREC = [0.46, 0.50, 0.52, 0.51, 0.48, 0.50, 0.49, 0.48, 0.52, 0.56, 0.50, 0.47]
FNIGHT_PET = [0.10] * 12
FNIGHT_TEMP = [-0.76, -1.30, -1.88, -2.38, -2.72, -2.75, -2.74, -3.04, -2.44, -1.60, -0.94, -0.53]
# mhm_parameter.nml: &routing1 defaults
ROUT1_PARAM = [0.325, 0.075, 2.0, 0.1, 0.3]


def process_matrix(soil_case=1, pet_case=-1, rout_case=1):
    """processMatrix (nProcesses=11, 3) as numpy (3, 11); only column 1 matters here"""
    pm = np.zeros((3, 11), dtype=np.int32)
    pm[0, :] = [1, 1, soil_case, 1, pet_case, 1, 1, rout_case, 1, 0, 0]
    return pm


def make_params(rng, n, nH=2, nLAI=12, nLC=2, pet_case=-1):
    U = rng.uniform
    P = {}
    fs = U(0.0, 0.6, (nLC, 1, n))
    fs[:, :, rng.random(n) < 0.3] = 0.0  # unsealed cells take the frac_sealed == 0 branch
    P["L1_fSealed"] = fs
    P["L1_alpha"] = U(0.05, 0.6, (nLC, 1, n))
    P["L1_degDayInc"] = U(0.1, 0.9, (nLC, 1, n))
    P["L1_degDayNoPre"] = U(0.5, 4.0, (nLC, 1, n))
    P["L1_degDayMax"] = P["L1_degDayNoPre"] + U(0.0, 4.0, (nLC, 1, n))
    fr = U(0.05, 1.0, (nLC, nH, n))
    P["L1_fRoots"] = fr / fr.sum(axis=1, keepdims=True)
    mi = U(0.0, 2.0, (1, nLAI, n))
    mi[:, :, rng.random(n) < 0.1] = 0.0  # bare cells: interc_max <= eps branch
    P["L1_maxInter"] = mi
    P["L1_karstLoss"] = U(0.8, 1.0, (1, 1, n))
    k0 = U(1.0, 10.0, (nLC, 1, n))
    k1 = np.maximum(U(2.0, 40.0, (nLC, 1, n)), k0)
    P["L1_kFastFlow"], P["L1_kSlowFlow"] = k0, k1
    P["L1_kBaseFlow"] = k1 + U(0.0, 1000.0, (nLC, 1, n))
    P["L1_kPerco"] = U(2.0, 60.0, (nLC, 1, n))
    sat = U(50.0, 400.0, (nLC, nH, n))
    fc = sat * U(0.3, 0.7, (nLC, nH, n))
    P["L1_soilMoistSat"], P["L1_soilMoistFC"] = sat, fc
    P["L1_wiltingPoint"] = fc * U(0.2, 0.5, (nLC, nH, n))
    P["L1_soilMoistExp"] = U(1.5, 6.0, (nLC, nH, n))
    P["L1_jarvis_thresh_c1"] = U(0.3, 0.7, (1, 1, n))
    P["L1_tempThresh"] = U(-2.0, 2.0, (nLC, 1, n))
    P["L1_unsatThresh"] = U(5.0, 150.0, (1, 1, n))
    st = U(0.0, 5.0, (1, 1, n))
    st[:, :, rng.random(n) < 0.05] = 0.0  # water_thresh_sealed <= eps: huge() branch
    P["L1_sealedThresh"] = st
    P["L1_petLAIcorFactor"] = U(0.7, 1.3, (nLC, nLAI, n))
    P["L1_fAsp"] = U(0.8, 1.2, (1, 1, n))
    P["L1_HarSamCoeff"] = U(0.002, 0.003, (1, 1, n))
    P["L1_PrieTayAlpha"] = U(1.0, 1.4, (1, nLAI, n))
    P["L1_aeroResist"] = U(20.0, 200.0, (nLC, nLAI, n))
    P["L1_surfResist"] = U(30.0, 250.0, (1, nLAI, n))
    P["latitude"] = U(35.0, 60.0, (1, 1, n))
    return {k: np.ascontiguousarray(v) for k, v in P.items()}


def make_forcing(rng, n, n_meteo, hourly=True, pet_case=-1, first_step=1):
    """forcing rows first_step .. first_step+n_meteo-1, numpy (n_meteo, n)"""
    F = {}
    steps = np.arange(first_step - 1, first_step - 1 + n_meteo)
    if hourly:
        h = (steps % 24)[:, None].astype(np.float64)
        doy = ((steps // 24) % 365 + 1)[:, None].astype(np.float64)
        wet = rng.random((n_meteo, n)) < 0.2
        F["pre"] = np.where(wet, rng.gamma(0.7, 1.6, (n_meteo, n)), 0.0)
        t = 10.0 + 12.0 * np.sin(2 * np.pi * doy / 365.0) + 4.0 * np.sin(2 * np.pi * h / 24.0)
        F["temp"] = np.clip(t + rng.normal(0.0, 2.0, (n_meteo, n)), -100.0, 100.0)
        pet = np.maximum(0.0, 0.15 * np.sin(np.pi * (h - 6.0) / 12.0)) * (
            1.0 + 0.5 * np.sin(2 * np.pi * (doy - 80.0) / 365.0))
        F["pet"] = np.clip(pet * rng.uniform(0.8, 1.2, (n_meteo, n)), 0.0, 1000.0)
    else:
        doy = (steps % 365 + 1)[:, None].astype(np.float64)
        wet = rng.random((n_meteo, n)) < 0.5
        F["pre"] = np.where(wet, rng.gamma(0.7, 8.0, (n_meteo, n)), 0.0)
        t = 8.0 + 12.0 * np.sin(2 * np.pi * (doy - 100.0) / 365.0)
        F["temp"] = np.clip(t + rng.normal(0.0, 3.0, (n_meteo, n)), -100.0, 100.0)
        F["pet"] = np.clip((2.5 + 2.0 * np.sin(2 * np.pi * (doy - 100.0) / 365.0))
                           * rng.uniform(0.7, 1.3, (n_meteo, n)), 0.0, 1000.0)
        if pet_case == 1:
            d = rng.uniform(-1.0, 12.0, (n_meteo, n))  # a few tmax < tmin rows on purpose
            F["tmin"] = F["temp"] - 0.5 * d
            F["tmax"] = F["temp"] + 0.5 * d
        if pet_case in (2, 3):
            F["netrad"] = rng.uniform(-20.0, 250.0, (n_meteo, n))
        if pet_case == 3:
            F["absvappress"] = rng.uniform(300.0, 2000.0, (n_meteo, n))
            F["windspeed"] = rng.uniform(0.5, 8.0, (n_meteo, n))
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in F.items()}


def default_states(n, nH, horizon_depth):
    """mHM/mo_init_states.f90:280-300"""
    S = {
        "L1_inter": np.zeros(n),
        "L1_snowPack": np.full(n, 15.0),
        "L1_sealSTW": np.zeros(n),
        "L1_unsatSTW": np.full(n, 10.0),
        "L1_satSTW": np.full(n, 75.0),
    }
    sm = np.zeros((nH, n))
    for i in range(nH - 1):
        sm[i] = (horizon_depth[i] if i == 0 else horizon_depth[i] - horizon_depth[i - 1]) * 0.25
    sm[nH - 1] = (1500.0 - (horizon_depth[nH - 2] if nH > 1 else 0.0)) * 0.25
    S["L1_soilMoist"] = sm
    return S


def scheidegger_network(rng, nx, ny, fill=0.85):
    """River network on an nx x ny grid with an irregular mask: every cell drains to one of
    its three neighbours in +x (dy in {-1,0,1}) chosen at random; cells draining out of the
    mask are outlets.  Returns nodes in raster order (1-based ids) and the link vectors the
    reference builds in L11_set_network_topology (links in ascending from-node order)."""
    yy, xx = np.mgrid[0:ny, 0:nx]
    # smooth irregular mask: ellipse with a noisy rim, tuned to ~fill
    r = ((xx - nx / 2.0) / (nx / 2.0)) ** 2 + ((yy - ny / 2.0) / (ny / 2.0)) ** 2
    noise = rng.normal(0.0, 0.05, (ny, nx))
    thr = np.quantile(r + noise, fill)
    mask = (r + noise) <= thr
    ids = np.full((ny, nx), -1, dtype=np.int64)
    ids[mask] = np.arange(1, mask.sum() + 1)
    dy = rng.integers(-1, 2, (ny, nx))
    ty = yy + dy
    tx = xx + 1
    inside = (tx < nx) & (ty >= 0) & (ty < ny)
    tgt = np.full((ny, nx), -1, dtype=np.int64)
    tgt[inside] = ids[ty[inside], tx[inside]]
    node = ids[mask]
    to = tgt[mask]
    is_link = to > 0
    fromN = node[is_link].astype(np.int32)
    toN = to[is_link].astype(np.int32)
    nNodes = int(mask.sum())
    return {"nNodes": nNodes, "nOutlets": nNodes - len(fromN), "fromN": fromN, "toN": toN,
            "mask": mask}


def make_network(rng, nx, ny, routing_order, fill=0.85, rout_case=1, l1_factor=1, nLC=2,
                 n_gauges=3, inflow=None):
    """network + L1<->L11 mapping.  l1_factor = 1: L11 == L1; 2: every L11 node holds up to
    4 L1 cells (map_flag true); -2: every L1 cell holds 4 L11 nodes (map_flag false)."""
    net = scheidegger_network(rng, nx, ny, fill)
    nNodes = net["nNodes"]
    nLinks = len(net["fromN"])
    pad = lambda a: np.concatenate([a, np.full(nNodes - len(a), -9999, dtype=np.int32)])
    rOrder, netPerm = routing_order(nNodes, net["fromN"], net["toN"])
    net["fromN"], net["toN"] = pad(net["fromN"]), pad(net["toN"])
    net["netPerm"], net["rOrder"] = netPerm, rOrder
    mask = net.pop("mask")
    ids = np.zeros(mask.shape, dtype=np.int32)
    ids[mask] = np.arange(1, nNodes + 1)
    if l1_factor == 1:
        nCells1 = nNodes
        net["L1_L11_Id"] = np.arange(1, nNodes + 1, dtype=np.int32)
        net["L11_L1_Id"] = np.arange(1, nNodes + 1, dtype=np.int32)
        net["map_flag"] = 1
    elif l1_factor > 1:
        f = l1_factor
        fine = np.kron(ids, np.ones((f, f), dtype=np.int32))
        keep = (fine > 0) & (rng.random(fine.shape) < 0.95)  # ragged: some L1 cells missing
        # every node keeps at least its first sub-cell
        first = np.zeros_like(fine, dtype=bool)
        first[::f, ::f] = ids > 0
        keep |= first
        net["L1_L11_Id"] = fine[keep].astype(np.int32)
        nCells1 = int(keep.sum())
        net["L11_L1_Id"] = np.arange(1, nNodes + 1, dtype=np.int32)  # unused for map_flag
        net["map_flag"] = 1
    else:
        f = -l1_factor
        cy, cx = (mask.shape[0] + f - 1) // f, (mask.shape[1] + f - 1) // f
        coarse_has = np.zeros((cy, cx), dtype=bool)
        ys, xs = np.nonzero(mask)
        coarse_has[ys // f, xs // f] = True
        cid = np.zeros((cy, cx), dtype=np.int32)
        cid[coarse_has] = np.arange(1, coarse_has.sum() + 1)
        nCells1 = int(coarse_has.sum())
        net["L11_L1_Id"] = cid[ys // f, xs // f].astype(np.int32)
        net["L1_L11_Id"] = np.ones(nCells1, dtype=np.int32)  # unused for !map_flag
        net["map_flag"] = 0
    net["nCells1"] = nCells1
    net["L1_areaCell"] = rng.uniform(0.9, 1.1, nCells1) * 16.0   # km2
    net["L11_areaCell"] = rng.uniform(0.9, 1.1, nNodes) * 16.0 * (
        l1_factor ** 2 if l1_factor > 1 else (1.0 / l1_factor ** 2 if l1_factor < 0 else 1.0))
    # gauges: the node collecting most links + random interior nodes
    acc = np.ones(nNodes + 1, dtype=np.int64)
    for k in range(nLinks):
        i = netPerm[k] - 1
        acc[net["toN"][i]] += acc[net["fromN"][i]]
    g = [int(np.argmax(acc[1:]) + 1)]
    while len(g) < min(n_gauges, nNodes):
        c = int(rng.integers(1, nNodes + 1))
        if c not in g:
            g.append(c)
    net["gaugeNodeList"] = np.array(g, dtype=np.int32)
    net["gaugeIndexList"] = np.arange(1, len(g) + 1, dtype=np.int32)
    net["nGaugesTotal"] = len(g)
    if inflow:
        # inflow gauges at nodes with a fair upstream area; (headwater flag per gauge)
        cand = np.argsort(-acc[1:])[5:5 + len(inflow)] + 1
        net["InflowGaugeNodeList"] = cand.astype(np.int32)
        net["InflowGaugeIndexList"] = np.arange(1, len(inflow) + 1, dtype=np.int32)
        net["InflowGaugeHeadwater"] = np.array([int(b) for b in inflow], dtype=np.int32)
        net["nInflowTotal"] = len(inflow)
    else:
        net["InflowGaugeNodeList"] = np.zeros(0, dtype=np.int32)
        net["InflowGaugeIndexList"] = np.zeros(0, dtype=np.int32)
        net["InflowGaugeHeadwater"] = np.zeros(0, dtype=np.int32)
        net["nInflowTotal"] = 0
    net["processCase"] = rout_case
    net["L11_length"] = rng.uniform(1000.0, 30000.0, nNodes)
    net["L11_slope"] = rng.uniform(0.001, 0.2, nNodes)
    net["L11_nLinkFracFPimp"] = rng.uniform(0.0, 0.3, (nLC, nNodes))
    net["rout_param"] = np.array(ROUT1_PARAM)
    return net


def make_problem(nx=20, ny=12, n_days=4, nH=2, nLAI=12, nLC=2, hourly=True, soil_case=1,
                 pet_case=-1, rout_case=1, l1_factor=1, routing=True, timestep_h=1, seed=SEED,
                 start=(1990, 12, 30), lc_switch_year=1991, read_weights=False, inflow=None,
                 celerity=1.5, fill=0.85, n_gauges=3, timeStep_LAI_input=0, routing_order=None):
    """A complete single-domain problem.  The default period straddles a year change so that
    the land-cover scene, the LAI month and evap_coeff all switch inside the run."""
    import datetime

    rng = np.random.default_rng(seed)
    prob = {"nH": nH, "nLAI": nLAI, "nLC": nLC, "timestep_h": timestep_h, "hourly": hourly,
            "soil_case": soil_case, "pet_case": pet_case, "rout_case": rout_case,
            "read_weights": read_weights}
    net = None
    if routing:
        if routing_order is None:
            from .interface import routing_order
        net = make_network(rng, nx, ny, routing_order, fill, rout_case, l1_factor, nLC, n_gauges,
                           inflow)
        n = net["nCells1"]
    else:
        n = int(nx * ny * fill)
    prob["nCells"] = n
    prob["net"] = net
    prob["processMatrix"] = process_matrix(soil_case, pet_case, rout_case if routing else 0)
    nTstepDay = 24 // timestep_h
    nT = n_days * nTstepDay
    d0 = datetime.date(*start)
    jul = JUL_1990_01_01 + (d0 - datetime.date(1990, 1, 1)).days
    years = list(range(start[0] - 1, start[0] + n_days // 365 + 3))
    prob["time"] = {
        "jul_start": jul, "nTimeSteps": nT, "warming_days": 0,
        "timeStep_LAI_input": timeStep_LAI_input, "lc_year_start": years[0],
        "LCyearId": np.array([1 if (y < lc_switch_year or nLC == 1) else min(2, nLC) for y in years],
                             dtype=np.int32),
    }
    prob["nTstepForcingDay"] = 24 if hourly else 1
    n_meteo = nT * timestep_h if hourly else n_days
    if hourly:
        assert timestep_h == 1
    prob["params"] = make_params(rng, n, nH, nLAI, nLC, pet_case)
    prob["forcing"] = make_forcing(rng, n, n_meteo, hourly, pet_case)
    if read_weights:
        w = rng.uniform(0.5, 1.5, (24, 12, n)) / 24.0
        prob["weights"] = {"pre": w, "pet": np.ascontiguousarray(w[::-1]),
                           "temp": np.ascontiguousarray(1.0 + 0.01 * (w - w.mean()))}
    prob["horizon_depth"] = np.array([200.0 * (i + 1) for i in range(nH)])
    prob["states0"] = default_states(n, nH, prob["horizon_depth"])
    if net is not None:
        nI = net["nInflowTotal"]
        prob["inflowQ"] = np.ascontiguousarray(rng.uniform(5.0, 50.0, (max(nI, 0), n_days)))
        if rout_case in (2, 3):
            net["celerity"] = celerity
    return prob
