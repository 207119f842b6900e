"""Host-side mirror of the reference's run interface over the C ABI.

The reference drives the hot path from `mhm_interface_run_*` (mHM/mo_mhm_interface_run.f90);
this module offers the same steps with the same names for Python callers (tests, bench.py):

    ctx = Context()                       # mhm_interface_init  -> mhm_cuda_init
    dom = ctx.register_domain(1, ...)     # mhm_initialize      -> mhm_cuda_register_domain
    dom.set_param("L1_fSealed", arr)      # mpr_eval / restart  -> mhm_cuda_set_param
    dom.do_time_step(tt, idx)             # mhm_interface_run_do_time_step (per-step seam)
    dom.run_steps(tt_first, n)            # TimeLoop body for a block of steps
    dom.get_variable("L1_soilMoist")      # pybind get%L1_variable

Array convention: numpy C-order arrays whose shape is the reversed Fortran shape, i.e.
Fortran (nCells, dim2, dim3) <-> numpy (dim3, dim2, nCells); memory is identical.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FLUX, METEO, MRM_STATE, PARAM, STATE, check

# reference variable name -> enum name
PARAM_NAMES = {
    "L1_fSealed": "MHM_P_FSEALED",
    "L1_alpha": "MHM_P_ALPHA",
    "L1_degDayInc": "MHM_P_DEGDAYINC",
    "L1_degDayMax": "MHM_P_DEGDAYMAX",
    "L1_degDayNoPre": "MHM_P_DEGDAYNOPRE",
    "L1_fRoots": "MHM_P_FROOTS",
    "L1_maxInter": "MHM_P_MAXINTER",
    "L1_karstLoss": "MHM_P_KARSTLOSS",
    "L1_kFastFlow": "MHM_P_KFASTFLOW",
    "L1_kSlowFlow": "MHM_P_KSLOWFLOW",
    "L1_kBaseFlow": "MHM_P_KBASEFLOW",
    "L1_kPerco": "MHM_P_KPERCO",
    "L1_soilMoistFC": "MHM_P_SOILMOISTFC",
    "L1_soilMoistSat": "MHM_P_SOILMOISTSAT",
    "L1_soilMoistExp": "MHM_P_SOILMOISTEXP",
    "L1_jarvis_thresh_c1": "MHM_P_JARVIS_C1",
    "L1_tempThresh": "MHM_P_TEMPTHRESH",
    "L1_unsatThresh": "MHM_P_UNSATTHRESH",
    "L1_sealedThresh": "MHM_P_SEALEDTHRESH",
    "L1_wiltingPoint": "MHM_P_WILTINGPOINT",
    "L1_petLAIcorFactor": "MHM_P_PETLAICORFACTOR",
    "L1_fAsp": "MHM_P_FASP",
    "L1_HarSamCoeff": "MHM_P_HARSAMCOEFF",
    "L1_PrieTayAlpha": "MHM_P_PRIETAYALPHA",
    "L1_aeroResist": "MHM_P_AERORESIST",
    "L1_surfResist": "MHM_P_SURFRESIST",
    "latitude": "MHM_P_LATITUDE",
}
STATE_NAMES = {
    "L1_inter": "MHM_S_INTER",
    "L1_snowPack": "MHM_S_SNOWPACK",
    "L1_sealSTW": "MHM_S_SEALSTW",
    "L1_unsatSTW": "MHM_S_UNSATSTW",
    "L1_satSTW": "MHM_S_SATSTW",
    "L1_soilMoist": "MHM_S_SOILMOIST",
}
FLUX_NAMES = {
    "L1_pet_calc": "MHM_F_PET_CALC",
    "L1_temp_calc": "MHM_F_TEMP_CALC",
    "L1_prec_calc": "MHM_F_PREC_CALC",
    "L1_aETCanopy": "MHM_F_AETCANOPY",
    "L1_aETSealed": "MHM_F_AETSEALED",
    "L1_baseflow": "MHM_F_BASEFLOW",
    "L1_fastRunoff": "MHM_F_FASTRUNOFF",
    "L1_melt": "MHM_F_MELT",
    "L1_percol": "MHM_F_PERCOL",
    "L1_preEffect": "MHM_F_PREEFFECT",
    "L1_rain": "MHM_F_RAIN",
    "L1_runoffSeal": "MHM_F_RUNOFFSEAL",
    "L1_slowRunoff": "MHM_F_SLOWRUNOFF",
    "L1_snow": "MHM_F_SNOW",
    "L1_Throughfall": "MHM_F_THROUGHFALL",
    "L1_total_runoff": "MHM_F_TOTAL_RUNOFF",
    "L1_degDay": "MHM_F_DEGDAY",
    "L1_aETSoil": "MHM_F_AETSOIL",
    "L1_infilSoil": "MHM_F_INFILSOIL",
}
METEO_NAMES = {
    "pre": "MHM_M_PRE",
    "temp": "MHM_M_TEMP",
    "pet": "MHM_M_PET",
    "tmin": "MHM_M_TMIN",
    "tmax": "MHM_M_TMAX",
    "netrad": "MHM_M_NETRAD",
    "absvappress": "MHM_M_ABSVAPPRESS",
    "windspeed": "MHM_M_WINDSPEED",
}
MRM_STATE_NAMES = {
    "L11_qOUT": "MRM_S_QOUT",
    "L11_qTIN": "MRM_S_QTIN",
    "L11_qTR": "MRM_S_QTR",
    "L11_qMod": "MRM_S_QMOD",
    "L11_C1": "MRM_S_C1",
    "L11_C2": "MRM_S_C2",
}


def _pd(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.shape)
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    if a is None:
        return C.POINTER(C.c_int32)()
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.shape)
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def routing_order(nNodes, fromN, toN, nLinks=None):
    """L11_routing_order in O(nLinks) (host only): returns (rOrder, netPerm), padded to nNodes."""
    L = _lib.load()
    fromN = np.ascontiguousarray(fromN, dtype=np.int32)
    toN = np.ascontiguousarray(toN, dtype=np.int32)
    nLinks = len(fromN) if nLinks is None else nLinks
    rOrder = np.full(nNodes, -9999, dtype=np.int32)
    netPerm = np.full(nNodes, -9999, dtype=np.int32)
    check(L.mrm_routing_order(nNodes, nLinks, _pi(fromN), _pi(toN), _pi(rOrder), _pi(netPerm)))
    return rOrder, netPerm


def shared_rows(n_steps, nranks, rank):
    """(first_row, n_rows) of a shared forcing chunk that `rank` copies from its host (host only)"""
    L = _lib.load()
    a, b = C.c_int64(), C.c_int64()
    check(L.mhm_cuda_meteo_shared_rows(n_steps, nranks, rank, C.byref(a), C.byref(b)))
    return a.value, b.value


def time_indices(time_cfg, timestep_h, nTstepForcingDay, tt_first, n_steps):
    """per-step calendar indices as the library derives them (host only)"""
    L = _lib.load()
    tc, keep = _time_config(time_cfg)
    out = (_lib.StepIndex * n_steps)()
    check(L.mhm_time_indices(C.byref(tc), timestep_h, nTstepForcingDay, tt_first, n_steps, out))
    del keep
    return out


def _time_config(cfg):
    tc = _lib.TimeConfig()
    lc = np.ascontiguousarray(cfg["LCyearId"], dtype=np.int32)
    tc.jul_start = cfg["jul_start"]
    tc.nTimeSteps = cfg["nTimeSteps"]
    tc.warming_days = cfg.get("warming_days", 0)
    tc.timeStep_LAI_input = cfg.get("timeStep_LAI_input", 0)
    tc.lc_year_start = cfg["lc_year_start"]
    tc.lc_nyears = len(lc)
    tc.LCyearId = _pi(lc)
    return tc, lc


class Context:
    """mhm_cuda_init / mhm_cuda_finalize"""

    def __init__(self, device=-1):
        self.L = _lib.load()
        self.h = C.c_void_p()
        check(self.L.mhm_cuda_init(device, C.byref(self.h)))
        self.domains = {}

    def register_domain(self, iDomain, nCells, nHorizons, nLAI, nLCscenes, processMatrix,
                        timestep_h=1, read_states=False, nMembers=1):
        cfg = _lib.DomainConfig()
        pm = np.ascontiguousarray(processMatrix, dtype=np.int32)  # numpy (3, nProcesses)
        cfg.nCells, cfg.nHorizons, cfg.nLAI, cfg.nLCscenes = nCells, nHorizons, nLAI, nLCscenes
        cfg.nMembers = nMembers
        cfg.nProcesses = pm.shape[1]
        cfg.timestep_h = timestep_h
        cfg.read_states = int(read_states)
        cfg.c2TSTu = float(timestep_h) / 24.0  # mo_startup.f90:168
        cfg.processMatrix = _pi(pm)
        check(self.L.mhm_cuda_register_domain(self.h, iDomain, C.byref(cfg)))
        dom = Domain(self, iDomain, nCells, nHorizons, nLAI, nLCscenes, nMembers, timestep_h)
        self.domains[iDomain] = dom
        return dom

    def synchronize(self):
        check(self.L.mhm_cuda_synchronize(self.h))

    def set_math_mode(self, mode):
        check(self.L.mhm_cuda_set_math_mode(self.h, {"strict": 0, "fast": 1}.get(mode, mode)))

    def event_record(self, slot):
        check(self.L.mhm_cuda_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double()
        check(self.L.mhm_cuda_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def kernel_stats(self, which):
        ms, n = C.c_double(), C.c_int64()
        check(self.L.mhm_cuda_kernel_stats(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def kernel_stats_reset(self, enable_timing=True):
        check(self.L.mhm_cuda_kernel_stats_reset(self.h, int(enable_timing)))

    def measure_dfma_peak(self):
        v = C.c_double()
        check(self.L.mhm_cuda_measure_dfma_peak(self.h, C.byref(v)))
        return v.value

    # ---- the GPUs of one box (comm.cu) ---------------------------------------------------
    def comm_init(self, dist=None):
        """library-owned NCCL communicator over the ranks of torch.distributed `dist` (the unique id
        travels through dist.broadcast_object_list, like MPI_Bcast in the Fortran driver)"""
        if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
            check(self.L.mhm_cuda_comm_init(self.h, 1, 0, None))
            return 1, 0
        world, rank = dist.get_world_size(), dist.get_rank()
        buf = C.create_string_buffer(128)
        if rank == 0:
            check(self.L.mhm_cuda_comm_unique_id(buf))
        box = [buf.raw]
        dist.broadcast_object_list(box, src=0)
        check(self.L.mhm_cuda_comm_init(self.h, world, rank, box[0]))
        return world, rank

    def comm_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(self.L.mhm_cuda_comm_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"nranks": a.value, "rank": b.value, "nccl_version": c.value}

    def finalize(self):
        if self.h:
            check(self.L.mhm_cuda_finalize(self.h))
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.finalize()


class Domain:
    def __init__(self, ctx, iDomain, nCells, nH, nLAI, nLC, nMembers, timestep_h):
        self.ctx, self.L, self.h, self.id = ctx, ctx.L, ctx.h, iDomain
        self.nCells, self.nH, self.nLAI, self.nLC = nCells, nH, nLAI, nLC
        self.nMembers, self.timestep_h = nMembers, timestep_h
        self.nNodes = 0
        self.nGaugesTotal = 0
        self.nTimeSteps = 0

    # ---- parameters / states / fluxes ------------------------------------------------
    # `ld` / `offset`: the array is the module-global one of ALL domains (leading dimension
    # nCellsTot, this domain's first cell at offset = s1 - 1), like the sections
    # L1_fSealed(s1:e1, 1, yId) the Fortran driver passes (mo_mhm_interface_run.f90:394-457)
    def _glob(self, arr, ld, offset, n=None):
        n = self.nCells if n is None else n
        ld = arr.shape[-1] if ld is None else ld
        assert arr.shape[-1] == ld and offset >= 0 and offset + n <= ld, (arr.shape, ld, offset, n)
        return ld

    def set_param(self, name, arr, member=0, ld=None, offset=0):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        if arr.ndim == 1:
            arr = arr.reshape(1, 1, -1)
        dim3, dim2, _ = arr.shape
        ld = self._glob(arr, ld, offset)
        check(self.L.mhm_cuda_set_param(self.h, self.id, member, PARAM[PARAM_NAMES[name]], _pd(arr),
                                        ld, offset, dim2, dim3))

    def get_param(self, name, dim2, dim3, member=0):
        out = np.zeros((dim3, dim2, self.nCells))
        check(self.L.mhm_cuda_get_param(self.h, self.id, member, PARAM[PARAM_NAMES[name]], _pd(out),
                                        self.nCells, 0, dim2, dim3))
        return out

    def set_state(self, name, arr, member=0, ld=None, offset=0):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        ld = self._glob(arr, ld, offset)
        check(self.L.mhm_cuda_set_state(self.h, self.id, member, STATE[STATE_NAMES[name]], _pd(arr),
                                        ld, offset))

    def get_state(self, name, member=0, out=None, offset=0):
        if out is None:
            out = np.zeros((self.nH, self.nCells)) if name == "L1_soilMoist" else np.zeros(self.nCells)
        ld = self._glob(out, None, offset)
        check(self.L.mhm_cuda_get_state(self.h, self.id, member, STATE[STATE_NAMES[name]], _pd(out),
                                        ld, offset))
        return out

    def get_flux(self, name, member=0, out=None, offset=0):
        two = name in ("L1_aETSoil", "L1_infilSoil")
        if out is None:
            out = np.zeros((self.nH, self.nCells)) if two else np.zeros(self.nCells)
        ld = self._glob(out, None, offset)
        check(self.L.mhm_cuda_get_flux(self.h, self.id, member, FLUX[FLUX_NAMES[name]], _pd(out),
                                       ld, offset))
        return out

    # host coherence (pybind get%L1_variable, restart writers): bind the module globals once,
    # sync_to_host then refreshes this domain's section of every bound array (member 0)
    def bind_host_state(self, name, arr, offset=0):
        assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
        ld = self._glob(arr, None, offset)
        check(self.L.mhm_cuda_bind_host_state(self.h, self.id, STATE[STATE_NAMES[name]], _pd(arr), ld, offset))

    def bind_host_flux(self, name, arr, offset=0):
        assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
        ld = self._glob(arr, None, offset)
        check(self.L.mhm_cuda_bind_host_flux(self.h, self.id, FLUX[FLUX_NAMES[name]], _pd(arr), ld, offset))

    def sync_to_host(self):
        check(self.L.mhm_cuda_sync_to_host(self.h, self.id))

    def get_variable(self, name, member=0):
        """pybind get%L1_variable equivalent"""
        if name in STATE_NAMES:
            return self.get_state(name, member)
        if name in FLUX_NAMES:
            return self.get_flux(name, member)
        if name in MRM_STATE_NAMES:
            return self.get_routing_state(name, member)
        raise KeyError(name)

    def states_default_init(self, HorizonDepth_mHM):
        d = np.ascontiguousarray(HorizonDepth_mHM, dtype=np.float64)
        check(self.L.mhm_cuda_states_default_init(self.h, self.id, _pd(d)))

    # ---- meteo / time ------------------------------------------------------------------
    def set_meteo_config(self, pet_case, nTstepForcingDay, is_hourly_forcing, read_meteo_weights,
                         fnight_prec, fnight_pet, fnight_temp, evap_coeff):
        mc = _lib.MeteoConfig()
        mc.pet_case, mc.nTstepForcingDay = pet_case, nTstepForcingDay
        mc.is_hourly_forcing, mc.read_meteo_weights = int(is_hourly_forcing), int(read_meteo_weights)
        for m in range(12):  # mo_meteo_handler.f90:404-406
            mc.fnight_prec[m] = fnight_prec[m]
            mc.fnight_pet[m] = fnight_pet[m]
            mc.fnight_temp[m] = fnight_temp[m]
            mc.fday_prec[m] = 1.0 - fnight_prec[m]
            mc.fday_pet[m] = 1.0 - fnight_pet[m]
            mc.fday_temp[m] = -1.0 * fnight_temp[m]
            mc.evap_coeff[m] = evap_coeff[m]
        check(self.L.mhm_cuda_set_meteo_config(self.h, self.id, C.byref(mc)))

    def set_meteo(self, var, arr, first_step=1, ld=None, offset=0):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.ndim == 2
        ld = self._glob(arr, ld, offset)
        check(self.L.mhm_cuda_set_meteo(self.h, self.id, METEO[METEO_NAMES[var]], _pd(arr),
                                        ld, offset, first_step, arr.shape[0]))

    def set_meteo_l2(self, var, data2, mask2, cellsize2, mask1, cellsize1, first_step=1):
        """level-2 chunk, numpy (n_steps, ncols2, nrows2) float64 or float32 == Fortran
        (nrows2, ncols2, n_steps); masks numpy (ncols, nrows) 0/1"""
        a = np.ascontiguousarray(data2)
        assert a.dtype in (np.float64, np.float32) and a.ndim == 3
        m2 = np.ascontiguousarray(mask2, dtype=np.int32)
        m1 = np.ascontiguousarray(mask1, dtype=np.int32)
        assert a.shape[1:] == m2.shape
        check(self.L.mhm_cuda_set_meteo_l2(self.h, self.id, METEO[METEO_NAMES[var]],
                                           C.c_void_p(a.ctypes.data), int(a.dtype == np.float32), m2.shape[1],
                                           m2.shape[0], _pi(m2), float(cellsize2), m1.shape[1], m1.shape[0], _pi(m1),
                                           float(cellsize1), first_step, a.shape[0]))

    def get_meteo(self, var, n_steps, first_step=1):
        out = np.zeros((n_steps, self.nCells))
        check(self.L.mhm_cuda_get_meteo(self.h, self.id, METEO[METEO_NAMES[var]], _pd(out), self.nCells,
                                        first_step, n_steps))
        return out

    def set_meteo_host_ptr(self, var, ptr, ld, first_step, n_steps, async_copy=False, offset=0):
        """upload from a raw host pointer (e.g. pinned torch tensor .data_ptr()); async_copy:
        return before the copy has finished (double buffered, own stream)"""
        fn = self.L.mhm_cuda_set_meteo_async if async_copy else self.L.mhm_cuda_set_meteo
        check(fn(self.h, self.id, METEO[METEO_NAMES[var]], C.cast(ptr, C.POINTER(C.c_double)), ld, offset,
                 first_step, n_steps))

    def set_meteo_shared(self, var, ptr, ld, first_step, n_steps, is_f32=False, offset=0):
        """forcing shared by all ranks of the context's communicator: this rank copies only its rows
        (shared_rows) from the host pointer, the rest arrives by NCCL all-gather"""
        check(self.L.mhm_cuda_set_meteo_shared(self.h, self.id, METEO[METEO_NAMES[var]], C.c_void_p(ptr),
                                               int(is_f32), ld, offset, first_step, n_steps))

    def meteo_h2d_bytes(self):
        b = C.c_int64()
        check(self.L.mhm_cuda_meteo_h2d_bytes(self.h, self.id, C.byref(b)))
        return b.value

    def set_meteo_device(self, var, dev_ptr, first_step, n_steps):
        check(self.L.mhm_cuda_set_meteo_device(self.h, self.id, METEO[METEO_NAMES[var]],
                                               C.c_void_p(dev_ptr), first_step, n_steps))

    def set_meteo_weights(self, var, arr, ld=None, offset=0):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape[:2] == (24, 12)
        ld = self._glob(arr, ld, offset)
        check(self.L.mhm_cuda_set_meteo_weights(self.h, self.id, METEO[METEO_NAMES[var]], _pd(arr),
                                                ld, offset))

    def set_time(self, time_cfg):
        tc, keep = _time_config(time_cfg)
        check(self.L.mhm_cuda_set_time(self.h, self.id, C.byref(tc)))
        self.nTimeSteps = time_cfg["nTimeSteps"]
        del keep

    # ---- stepping ----------------------------------------------------------------------
    def do_time_step(self, tt, idx):
        """per-step seam (B1): idx is one mhm_step_index"""
        check(self.L.mhm_cuda_cell_step(self.h, self.id, tt, C.byref(idx)))

    def run_steps(self, tt_first, n_steps):
        check(self.L.mhm_cuda_run_steps(self.h, self.id, tt_first, n_steps))

    def set_outputs(self, outputFlxState, timeStep_model_outputs):
        """mhm_outputs.nml: outputFlxState(1:21), timeStep_model_outputs"""
        f = np.ascontiguousarray(outputFlxState, dtype=np.int32)
        assert f.shape == (21,)
        check(self.L.mhm_cuda_set_outputs(self.h, self.id, _pi(f), int(timeStep_model_outputs)))

    def output_windows(self):
        """model steps that closed the output windows of the last run_steps call"""
        n = C.c_int32()
        check(self.L.mhm_cuda_get_output_windows(self.h, self.id, C.byref(n), C.POINTER(C.c_int32)(), 0))
        tt = np.zeros(max(1, n.value), dtype=np.int32)
        check(self.L.mhm_cuda_get_output_windows(self.h, self.id, C.byref(n), _pi(tt), n.value))
        return tt[: n.value].tolist()

    def get_output(self, window, variable, horizon=0, member=0):
        out = np.zeros(self.nCells)
        check(self.L.mhm_cuda_get_output(self.h, self.id, member, window, variable, horizon, _pd(out)))
        return out

    def set_optisim(self, sm=None, et=None, tws=None, bfi=False):
        """calibration aggregates of mhm_interface_run_update_optisim: sm = (timeStepInput, nTime,
        nSoilHorizons_sm_input), et / tws = (timeStepInput, nTime); bfi switches the BFI sums on"""
        c = _lib.OptisimConfig()
        if sm is not None:
            c.sm_on, c.sm_timeStepInput, c.sm_nTime, c.nSoilHorizons_sm_input = 1, int(sm[0]), int(sm[1]), int(sm[2])
        if et is not None:
            c.et_on, c.et_timeStepInput, c.et_nTime = 1, int(et[0]), int(et[1])
        if tws is not None:
            c.tws_on, c.tws_timeStepInput, c.tws_nTime = 1, int(tws[0]), int(tws[1])
        c.bfi_on = int(bool(bfi))
        self._opt_ntime = {"sm": c.sm_nTime, "et": c.et_nTime, "tws": c.tws_nTime}
        check(self.L.mhm_cuda_set_optisim(self.h, self.id, C.byref(c)))

    def get_optisim(self, which, member=0):
        """dataSim of 'sm' | 'et' | 'tws' as (nTime, nCells) = Fortran (nCells, nTime)"""
        out = np.zeros((self._opt_ntime[which], self.nCells))
        check(self.L.mhm_cuda_get_optisim(self.h, self.id, member, ("sm", "et", "tws").index(which), _pd(out),
                                          self.nCells, 0))
        return out

    def get_bfi_sums(self, cell_area, member=0):
        """(BFI_qBF_sum, BFI_qT_sum) of mo_mhm_interface_run.f90:630-636"""
        a = np.ascontiguousarray(cell_area, dtype=np.float64)
        qb, qt = C.c_double(), C.c_double()
        check(self.L.mhm_cuda_get_bfi_sums(self.h, self.id, member, _pd(a), C.byref(qb), C.byref(qt)))
        return qb.value, qt.value

    def keep_runoff_history(self, keep=True):
        check(self.L.mhm_cuda_keep_runoff_history(self.h, self.id, int(keep)))

    def get_runoff_history(self, n_steps, member=0):
        out = np.zeros((n_steps, self.nCells))
        check(self.L.mhm_cuda_get_runoff_history(self.h, self.id, member, _pd(out), self.nCells))
        return out

    # ---- routing -------------------------------------------------------------------------
    def set_network(self, net):
        nw = _lib.Network()
        keep = []

        def ip(key):
            a = net.get(key)
            if a is None:
                return _pi(None)
            a = np.ascontiguousarray(a, dtype=np.int32)
            keep.append(a)
            return _pi(a)

        def dp(key):
            a = np.ascontiguousarray(net[key], dtype=np.float64)
            keep.append(a)
            return _pd(a)

        nw.nNodes, nw.nOutlets, nw.map_flag = net["nNodes"], net["nOutlets"], int(net["map_flag"])
        nw.nGauges = len(net.get("gaugeNodeList", []))
        nw.nInflowGauges = len(net.get("InflowGaugeNodeList", []))
        nw.nGaugesTotal = net.get("nGaugesTotal", nw.nGauges)
        nw.nInflowTotal = net.get("nInflowTotal", nw.nInflowGauges)
        nw.processCase = net["processCase"]
        for k in ("L1_L11_Id", "L11_L1_Id", "netPerm", "fromN", "toN", "gaugeIndexList",
                  "gaugeNodeList", "InflowGaugeIndexList", "InflowGaugeHeadwater",
                  "InflowGaugeNodeList"):
            setattr(nw, k, ip(k))
        nw.L1_areaCell = dp("L1_areaCell")
        nw.L11_areaCell = dp("L11_areaCell")
        # sub-catchment sharding (mhm_b200.shard.extract)
        nw.nGhostSources = len(net.get("ghostSourceNodeList", []))
        nw.nExports = len(net.get("exportNodeList", []))
        nw.ghostSourceNodeList = ip("ghostSourceNodeList")
        nw.exportNodeList = ip("exportNodeList")
        nw.lastSinkNode = int(net.get("lastSinkNode", 0))
        nw.ssMax = float(net.get("ssMax", 0.0)) if (nw.nGhostSources or nw.nExports or "ghostSourceNodeList" in net) else 0.0
        check(self.L.mrm_cuda_set_network(self.h, self.id, C.byref(nw)))
        self.nNodes, self.nGaugesTotal = nw.nNodes, nw.nGaugesTotal
        del keep

    def set_reg_rout(self, param5, L11_length, L11_slope, L11_nLinkFracFPimp, member=0):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (param5, L11_length, L11_slope,
                                                                  L11_nLinkFracFPimp)]
        check(self.L.mrm_cuda_set_reg_rout(self.h, self.id, member, *[_pd(x) for x in a]))

    def set_c1c2(self, C1, C2, TSrout=0.0, member=0):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (C1, C2)]
        check(self.L.mrm_cuda_set_c1c2(self.h, self.id, member, _pd(a[0]), _pd(a[1]), TSrout))

    def set_inflow(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)  # numpy (nInflowTotal, nDays)
        check(self.L.mrm_cuda_set_inflow(self.h, self.id, _pd(Q), Q.shape[1]))

    def set_routing_state(self, name, arr, member=0, offset=0):
        """L11_* module globals: (nNodesTot) or (nNodesTot, 2); offset = s11 - 1"""
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        ld = self._glob(arr, None, offset, self.nNodes)
        check(self.L.mrm_cuda_set_state(self.h, self.id, member, MRM_STATE[MRM_STATE_NAMES[name]],
                                        _pd(arr), ld, offset))

    def get_routing_state(self, name, member=0, out=None, offset=0):
        two = name in ("L11_qTIN", "L11_qTR")
        if out is None:
            out = np.zeros((2, self.nNodes)) if two else np.zeros(self.nNodes)
        ld = self._glob(out, None, offset, self.nNodes)
        check(self.L.mrm_cuda_get_state(self.h, self.id, member, MRM_STATE[MRM_STATE_NAMES[name]],
                                        _pd(out), ld, offset))
        return out

    def route(self, tt, yId, timestep_rout, tsRoutFactorIn, RunToRout=None, InflowDischarge=None):
        """per-step seam (B3): one mRM_routing call"""
        r = None if RunToRout is None else np.ascontiguousarray(RunToRout, dtype=np.float64)
        q = None if InflowDischarge is None else np.ascontiguousarray(InflowDischarge, dtype=np.float64)
        null = C.POINTER(C.c_double)()
        check(self.L.mrm_cuda_route(self.h, self.id, 0, tt, yId, null if r is None else _pd(r),
                                    timestep_rout, tsRoutFactorIn, null if q is None else _pd(q)))

    def get_runoff(self, tt_first=1, n_steps=None, member=0, out=None):
        """mRM_runoff(tt, gauge) as numpy (nGaugesTotal, nTimeSteps)"""
        n_steps = self.nTimeSteps - tt_first + 1 if n_steps is None else n_steps
        if out is None:
            out = np.zeros((self.nGaugesTotal, self.nTimeSteps))
        check(self.L.mrm_cuda_get_runoff(self.h, self.id, member, _pd(out), self.nTimeSteps,
                                         tt_first, n_steps))
        return out
