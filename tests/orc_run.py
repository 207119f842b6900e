"""Run the CPU oracle on a problem dict (mhm_b200.synth layout).  Test infrastructure."""
import ctypes as C

import numpy as np

import orc
from mhm_b200 import synth

FLUX_ORDER = ["L1_pet_calc", "L1_temp_calc", "L1_prec_calc", "L1_aETCanopy", "L1_aETSealed",
              "L1_baseflow", "L1_fastRunoff", "L1_melt", "L1_percol", "L1_preEffect", "L1_rain",
              "L1_runoffSeal", "L1_slowRunoff", "L1_snow", "L1_Throughfall", "L1_total_runoff",
              "L1_degDay"]
STATE_ORDER = ["L1_inter", "L1_snowPack", "L1_sealSTW", "L1_unsatSTW", "L1_satSTW"]
P2FIELD = {
    "L1_fSealed": "fSealed", "L1_alpha": "alpha", "L1_degDayInc": "degDayInc",
    "L1_degDayMax": "degDayMax", "L1_degDayNoPre": "degDayNoPre", "L1_fRoots": "fRoots",
    "L1_maxInter": "maxInter", "L1_karstLoss": "karstLoss", "L1_kFastFlow": "kFastFlow",
    "L1_kSlowFlow": "kSlowFlow", "L1_kBaseFlow": "kBaseFlow", "L1_kPerco": "kPerco",
    "L1_soilMoistFC": "soilMoistFC", "L1_soilMoistSat": "soilMoistSat",
    "L1_soilMoistExp": "soilMoistExp", "L1_jarvis_thresh_c1": "jarvis_thresh_c1",
    "L1_tempThresh": "tempThresh", "L1_unsatThresh": "unsatThresh",
    "L1_sealedThresh": "sealedThresh", "L1_wiltingPoint": "wiltingPoint",
    "L1_petLAIcorFactor": "petLAIcorFactor", "L1_fAsp": "fAsp", "L1_HarSamCoeff": "HarSamCoeff",
    "L1_PrieTayAlpha": "PrieTayAlpha", "L1_aeroResist": "aeroResist",
    "L1_surfResist": "surfResist", "latitude": "latitude",
}
F2FIELD = {
    "L1_pet_calc": "pet_calc", "L1_temp_calc": "temp_calc", "L1_prec_calc": "prec_calc",
    "L1_aETCanopy": "aETCanopy", "L1_aETSealed": "aETSealed", "L1_baseflow": "baseflow",
    "L1_fastRunoff": "fastRunoff", "L1_melt": "melt", "L1_percol": "percol",
    "L1_preEffect": "preEffect", "L1_rain": "rain", "L1_runoffSeal": "runoffSeal",
    "L1_slowRunoff": "slowRunoff", "L1_snow": "snow", "L1_Throughfall": "throughfall",
    "L1_total_runoff": "total_runoff", "L1_degDay": "degDay", "L1_aETSoil": "aETSoil",
    "L1_infilSoil": "infilSoil",
}
S2FIELD = {"L1_inter": "inter", "L1_snowPack": "snowPack", "L1_sealSTW": "sealSTW",
           "L1_unsatSTW": "unsatSTW", "L1_satSTW": "satSTW", "L1_soilMoist": "soilMoist"}


class OracleRun:
    """Owns the numpy buffers behind an orc_domain and exposes results by reference name."""

    def __init__(self, prob, params=None, history=False, num_threads=1, outputs=None, max_windows=64,
                 optisim=None, bfi=False, cell_area=None, read_states=False):
        """outputs = (outputFlxState[21], timeStep_model_outputs) switches the gridded output
        accumulation on; results in self.out_windows() after run().
        optisim = {"sm": (timeStepInput, nTime, nSoilHorizons_sm_input), "et": (timeStepInput, nTime),
        "tws": (timeStepInput, nTime)} switches the calibration aggregates on (self.opt[...] =
        dataSim as (nTime, nCells)); bfi=True the BFI sums (d.bfi_qBF_sum / d.bfi_qT_sum)."""
        self.prob = prob
        self.keep = []
        d = self.d = orc.OrcDomain()
        n, nH = prob["nCells"], prob["nH"]
        d.nCells, d.nH, d.nLAI, d.nLC = n, nH, prob["nLAI"], prob["nLC"]
        d.pc_soil, d.pc_pet = prob["soil_case"], prob["pet_case"]
        d.read_states = int(read_states)  # mo_mhm.f90:448-450, mo_mrm_routing.f90:211
        d.timestep_h = prob["timestep_h"]
        d.nTstepDay = 24 // prob["timestep_h"]
        t = prob["time"]
        d.jul_start, d.nTimeSteps = t["jul_start"], t["nTimeSteps"]
        d.warming_days, d.timeStep_LAI_input = t["warming_days"], t["timeStep_LAI_input"]
        d.lc_year_start, d.lc_nyears = t["lc_year_start"], len(t["LCyearId"])
        d.LCyearId = self._i(t["LCyearId"])
        d.nTstepForcingDay = prob["nTstepForcingDay"]
        d.is_hourly_forcing = int(prob["hourly"])
        d.read_meteo_weights = int(prob["read_weights"])
        F = prob["forcing"]
        d.nMeteoSteps = F["pre"].shape[0]
        for var, fld in (("pre", "pre"), ("temp", "temp"), ("pet", "pet"), ("tmin", "tmin"),
                         ("tmax", "tmax"), ("netrad", "netrad"), ("absvappress", "absvappress"),
                         ("windspeed", "windspeed")):
            if var in F:
                setattr(d, fld, self._d(F[var]))
        if prob["read_weights"]:
            d.pre_weights = self._d(prob["weights"]["pre"])
            d.temp_weights = self._d(prob["weights"]["temp"])
            d.pet_weights = self._d(prob["weights"]["pet"])
        for m in range(12):
            d.fnight_prec[m] = synth.FNIGHT_PREC[m]
            d.fnight_pet[m] = synth.FNIGHT_PET[m]
            d.fnight_temp[m] = synth.FNIGHT_TEMP[m]
            d.fday_prec[m] = 1.0 - synth.FNIGHT_PREC[m]
            d.fday_pet[m] = 1.0 - synth.FNIGHT_PET[m]
            d.fday_temp[m] = -1.0 * synth.FNIGHT_TEMP[m]
            d.evap_coeff[m] = synth.EVAP_COEFF[m]
        d.c2TSTu = prob["timestep_h"] / 24.0
        P = prob["params"] if params is None else params
        for name, fld in P2FIELD.items():
            if name in P:
                setattr(d, fld, self._d(P[name]))
        self.S = {k: np.array(v, dtype=np.float64, copy=True) for k, v in prob["states0"].items()}
        for name, fld in S2FIELD.items():
            setattr(d, fld, self._d(self.S[name]))
        self.F = {}
        for name, fld in F2FIELD.items():
            two = name in ("L1_aETSoil", "L1_infilSoil")
            self.F[name] = np.zeros((nH, n)) if two else np.zeros(n)
            setattr(d, fld, self._d(self.F[name]))
        net = prob.get("net")
        d.do_routing = int(net is not None)
        d.pc_rout = prob["rout_case"] if net is not None else 0
        nT = t["nTimeSteps"]
        if net is not None:
            nn = net["nNodes"]
            d.nNodes, d.nOutlets, d.map_flag = nn, net["nOutlets"], int(net["map_flag"])
            d.nGauges = len(net["gaugeNodeList"])
            d.nInflowGauges = len(net["InflowGaugeNodeList"])
            d.nGaugesTotal, d.nInflowTotal = net["nGaugesTotal"], net["nInflowTotal"]
            d.L1_areaCell, d.L11_areaCell = self._d(net["L1_areaCell"]), self._d(net["L11_areaCell"])
            for k in ("L1_L11_Id", "L11_L1_Id", "netPerm", "fromN", "toN", "gaugeIndexList",
                      "gaugeNodeList", "InflowGaugeIndexList", "InflowGaugeHeadwater",
                      "InflowGaugeNodeList"):
                setattr(d, k, self._i(net[k]))
            d.nDays = nT // d.nTstepDay
            d.InflowQ = self._d(prob["inflowQ"]) if net["nInflowTotal"] else self._d(np.zeros(1))
            d.L11_length, d.L11_slope = self._d(net["L11_length"]), self._d(net["L11_slope"])
            d.L11_nLinkFracFPimp = self._d(net["L11_nLinkFracFPimp"])
            rp = net["rout_param"] if params is None else params.get("rout_param", net["rout_param"])
            for i in range(5):
                d.rout_param[i] = rp[i]
            self.R = {"L11_C1": np.zeros(nn), "L11_C2": np.zeros(nn), "L11_qOUT": np.zeros(nn),
                      "L11_qTIN": np.zeros((2, nn)), "L11_qTR": np.zeros((2, nn)),
                      "L11_qMod": np.zeros(nn)}
            if prob["rout_case"] in (2, 3):
                self.R["L11_C1"][:] = net["C1"]
                self.R["L11_C2"][:] = net["C2"]
                d.L11_TSrout = net["TSrout"]
            for k, v in self.R.items():
                setattr(d, k, self._d(v))
            self.RunToRout = np.zeros(n)
            self.InflowDischarge = np.zeros(max(1, net["nInflowTotal"]))
            d.RunToRout, d.InflowDischarge = self._d(self.RunToRout), self._d(self.InflowDischarge)
            self.mRM_runoff = np.zeros((max(1, net["nGaugesTotal"]), nT))
            d.mRM_runoff = self._d(self.mRM_runoff)
        self.rs = orc.lib().orc_flux_record_size(nH)
        self.history = None
        if history:
            self.history = np.zeros((nT, self.rs, n))
            d.flux_history = self._d(self.history)
        d.num_threads = num_threads
        self.n_slots = 0
        if outputs is not None:
            flags, ts = outputs
            for i in range(21):
                d.out_flags[i] = int(flags[i])
            d.timeStep_model_outputs = int(ts)
            self.slot_var = np.zeros(64, dtype=np.int32)
            self.slot_hor = np.zeros(64, dtype=np.int32)
            self.slot_avg = np.zeros(64, dtype=np.int32)
            L = orc.lib()
            L.orc_output_slots.restype = C.c_int32
            L.orc_output_slots.argtypes = [C.POINTER(C.c_int32)] * 1 + [C.c_int32] + [C.POINTER(C.c_int32)] * 3
            fl = np.ascontiguousarray(flags, dtype=np.int32)
            self.n_slots = L.orc_output_slots(orc.iptr(fl), nH, orc.iptr(self.slot_var), orc.iptr(self.slot_hor),
                                              orc.iptr(self.slot_avg))
            self.out_acc = np.zeros((max(1, self.n_slots), n))
            self.out_win = np.zeros((max_windows, max(1, self.n_slots), n))
            self.out_win_tt = np.zeros(max_windows, dtype=np.int32)
            d.out_acc, d.out_win = self._d(self.out_acc), self._d(self.out_win)
            self.out_acc, self.out_win = self.keep[-2], self.keep[-1]
            d.out_win_tt = self._i(self.out_win_tt)
            self.out_win_tt = self.keep[-1]
            d.out_max_windows = max_windows

        self.opt = {}
        for w, key in enumerate(("sm", "et", "tws")):
            d.opt_avg_ts[w] = 1  # optidata_sim%init
            if optisim and key in optisim:
                cfg = optisim[key]
                d.opt_on[w], d.opt_timestep[w], d.opt_ntime[w] = 1, int(cfg[0]), int(cfg[1])
                if key == "sm":
                    d.opt_nhor_sm = int(cfg[2])
                ptr = self._d(np.zeros((int(cfg[1]), n)))
                self.opt[key] = self.keep[-1]
                setattr(d, "opt_" + key, ptr)
        if bfi:
            d.bfi_on = 1
            if net is None:
                d.L1_areaCell = self._d(cell_area)

    def _d(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.keep.append(a)
        return orc.dptr(a)

    def _i(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        self.keep.append(a)
        return orc.iptr(a)

    def run(self, tt_first, tt_last):
        rc = orc.lib().orc_run(C.byref(self.d), tt_first, tt_last)
        assert rc == 0

    def out_windows(self):
        """list of (tt_end, {(var, horizon): values}) of the written output windows"""
        res = []
        for w in range(self.d.out_nwin):
            fields = {(int(self.slot_var[sl]), int(self.slot_hor[sl])): self.out_win[w, sl]
                      for sl in range(self.n_slots)}
            res.append((int(self.out_win_tt[w]), fields))
        return res

    def hist(self, name, tt):
        """value of a flux/state after step tt (1-based) from the history"""
        nH = self.prob["nH"]
        H = self.history[tt - 1]
        if name in FLUX_ORDER:
            return H[FLUX_ORDER.index(name)]
        if name in STATE_ORDER:
            return H[17 + STATE_ORDER.index(name)]
        base = {"L1_aETSoil": 22, "L1_infilSoil": 22 + nH, "L1_soilMoist": 22 + 2 * nH}[name]
        return H[base: base + nH]

    def time_indices(self, n):
        arrs = [np.zeros(n, dtype=np.int32) for _ in range(8)]
        orc.lib().orc_time_indices(C.byref(self.d), n, *[orc.iptr(a) for a in arrs])
        keys = ["month", "hour", "yId", "iLAI", "iMeteoTS", "isday", "doy", "year"]
        return dict(zip(keys, arrs))


def case23_params(net):
    """fill net['C1'], net['C2'], net['TSrout'] like mrm_update_param (case 2: constant celerity)"""
    L = orc.lib()
    nn = net["nNodes"]
    C1, C2 = np.zeros(nn), np.zeros(nn)
    ts = L.orc_mrm_update_param_case2(nn, net["nOutlets"], orc.dptr(np.ascontiguousarray(net["L11_length"])),
                                      float(net["celerity"]), orc.dptr(C1), orc.dptr(C2))
    net["C1"], net["C2"], net["TSrout"] = C1, C2, ts
    return net


def case3_params(net, z0):
    """routing case 3 for a golden network: L0_streamNet, floored link lengths, L11_celerity from
    the L0 slopes (z0 = tests/golden/test_domain_l0.npz), then C1 / C2 / TSrout like
    mrm_update_param.  Everything through the oracle's restatements."""
    L = orc.lib()
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    mask0 = i32(z0["mask0"])                     # numpy (ncols0, nrows0) == Fortran (nrows0, ncols0)
    ncols0, nrows0 = mask0.shape
    fdir2 = np.full(mask0.shape, -9999, dtype=np.int32)
    fdir2[mask0 != 0] = z0["fDir0"]
    nn, nl = net["nNodes"], net["nNodes"] - net["nOutlets"]
    loc = [i32(net[k]) for k in ("netPerm", "fRow", "fCol", "tRow", "tCol")]
    sn2 = np.zeros(mask0.shape, dtype=np.int32)
    L.orc_stream_net(nrows0, ncols0, orc.iptr(fdir2), nl, *[orc.iptr(a) for a in loc], orc.iptr(sn2))
    stream = np.ascontiguousarray(sn2[mask0 != 0])
    length = np.array(net["L11_length"], dtype=np.float64)
    L.orc_length_floor(nn, orc.dptr(length))
    cel = np.zeros(nn)
    slope0 = np.ascontiguousarray(z0["slope0"], dtype=np.float64)
    L.orc_calc_celerity(nrows0, ncols0, orc.iptr(mask0), orc.iptr(fdir2), orc.iptr(stream), orc.dptr(slope0), nn, nl,
                        *[orc.iptr(a) for a in loc], float(net["slope_factor"]), orc.dptr(cel))
    C1, C2 = np.zeros(nn), np.zeros(nn)
    ts = L.orc_mrm_update_param_case3(nn, net["nOutlets"], orc.dptr(length), orc.dptr(cel), orc.dptr(C1), orc.dptr(C2))
    net.update(C1=C1, C2=C2, TSrout=ts, L11_length=length, L11_celerity=cel, streamNet0=stream)
    return net
