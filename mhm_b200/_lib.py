"""ctypes binding of libmhm_cuda.so (the C ABI of include/mhm_cuda.h).

There is no fallback: if the shared library is missing, or no CUDA device is present when
a context is created, the call raises.
"""
import ctypes as C
import os

from . import _cstruct

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
HEADER = os.path.join(ROOT, "include", "mhm_cuda.h")
LIBPATH = os.environ.get("MHM_CUDA_LIB", os.path.join(PKG, "libmhm_cuda.so"))


class MhmCudaError(RuntimeError):
    pass


class DomainConfig(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mhm_domain_config")


class MeteoConfig(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mhm_meteo_config")


class TimeConfig(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mhm_time_config")


class StepIndex(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mhm_step_index")


class Network(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mrm_network")


class MprL0Inputs(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mpr_l0_inputs")


class OptisimConfig(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mhm_optisim_config")


class MprSoilDb(C.Structure):
    _fields_ = _cstruct.parse_struct(HEADER, "mpr_soil_db")


PARAM = _cstruct.parse_enum(HEADER, "mhm_param_id")
STATE = _cstruct.parse_enum(HEADER, "mhm_state_id")
FLUX = _cstruct.parse_enum(HEADER, "mhm_flux_id")
METEO = _cstruct.parse_enum(HEADER, "mhm_meteo_var")
MRM_STATE = _cstruct.parse_enum(HEADER, "mrm_state_id")

_lib = None


def load():
    """Load libmhm_cuda.so and declare every prototype; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise MhmCudaError(
            "%s not found: build it with `make -C mhm_b200/csrc` (or __graft_entry__.build()); "
            "there is no CPU fallback" % LIBPATH
        )
    L = C.CDLL(LIBPATH)
    vp, i32, i64, d = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    sig = {
        "mhm_cuda_init": [C.c_int, C.POINTER(vp)],
        "mhm_cuda_finalize": [vp],
        "mhm_cuda_register_domain": [vp, i32, C.POINTER(DomainConfig)],
        "mhm_cuda_unregister_domain": [vp, i32],
        "mhm_cuda_set_param": [vp, i32, i32, i32, pd, i64, i64, i32, i32],
        "mhm_cuda_set_state": [vp, i32, i32, i32, pd, i64, i64],
        "mhm_cuda_get_state": [vp, i32, i32, i32, pd, i64, i64],
        "mhm_cuda_states_default_init": [vp, i32, pd],
        "mhm_cuda_get_flux": [vp, i32, i32, i32, pd, i64, i64],
        "mhm_cuda_set_meteo_config": [vp, i32, C.POINTER(MeteoConfig)],
        "mhm_cuda_set_meteo": [vp, i32, i32, pd, i64, i64, i64, i64],
        "mhm_cuda_set_meteo_async": [vp, i32, i32, pd, i64, i64, i64, i64],
        "mhm_cuda_set_meteo_device": [vp, i32, i32, vp, i64, i64],
        "mhm_cuda_comm_unique_id": [C.c_char_p],
        "mhm_cuda_comm_init": [vp, i32, i32, C.c_char_p],
        "mhm_cuda_comm_finalize": [vp],
        "mhm_cuda_comm_info": [vp, pi, pi, pi],
        "mhm_cuda_meteo_shared_rows": [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)],
        "mhm_cuda_set_meteo_shared": [vp, i32, i32, vp, i32, i64, i64, i64, i64],
        "mhm_cuda_meteo_h2d_bytes": [vp, i32, C.POINTER(i64)],
        "mhm_cuda_set_meteo_weights": [vp, i32, i32, pd, i64, i64],
        "mhm_cuda_set_time": [vp, i32, C.POINTER(TimeConfig)],
        "mhm_time_indices": [C.POINTER(TimeConfig), i32, i32, i32, i32, C.POINTER(StepIndex)],
        "mhm_cuda_cell_step": [vp, i32, i32, C.POINTER(StepIndex)],
        "mhm_cuda_run_steps": [vp, i32, i32, i32],
        "mhm_cuda_set_math_mode": [vp, i32],
        "mhm_cuda_bind_host_state": [vp, i32, i32, pd, i64, i64],
        "mhm_cuda_bind_host_flux": [vp, i32, i32, pd, i64, i64],
        "mhm_cuda_sync_to_host": [vp, i32],
        "mhm_cuda_get_runoff_history": [vp, i32, i32, pd, i64],
        "mhm_cuda_keep_runoff_history": [vp, i32, i32],
        "mhm_cuda_set_outputs": [vp, i32, pi, i32],
        "mhm_cuda_get_meteo": [vp, i32, i32, pd, i64, i64, i64],
        "mhm_cuda_set_meteo_l2": [vp, i32, i32, vp, i32, i32, i32, pi, d, i32, i32, pi, d, i64, i64],
        "mrm_partition_subcatchments": [i32, i32, pi, pi, pi, i32, pi],
        "mrm_cuda_set_deferred": [vp, i32, i32],
        "mrm_cuda_set_exchange": [vp, i32, pi, pi],
        "mrm_cuda_shard_run_steps": [vp, i32, i32, i32],
        "mrm_cuda_shard_flush": [vp, i32],
        "mrm_cuda_route_pending": [vp, i32],
        "mrm_cuda_export_outflow": [vp, i32, vp, i32],
        "mrm_cuda_import_outflow": [vp, i32, vp, i32],
        "mhm_cuda_get_output_windows": [vp, i32, pi, pi, i32],
        "mhm_cuda_set_optisim": [vp, i32, C.POINTER(OptisimConfig)],
        "mhm_cuda_get_optisim": [vp, i32, i32, i32, pd, i64, i64],
        "mhm_cuda_get_bfi_sums": [vp, i32, i32, pd, pd, pd],
        "mhm_cuda_get_output": [vp, i32, i32, i32, i32, i32, pd],
        "mrm_cuda_set_network": [vp, i32, C.POINTER(Network)],
        "mrm_routing_order": [i32, i32, pi, pi, pi, pi],
        "mrm_cuda_set_reg_rout": [vp, i32, i32, pd, pd, pd, pd],
        "mrm_cuda_set_c1c2": [vp, i32, i32, pd, pd, d],
        "mrm_cuda_set_state": [vp, i32, i32, i32, pd, i64, i64],
        "mrm_cuda_get_state": [vp, i32, i32, i32, pd, i64, i64],
        "mrm_cuda_set_inflow": [vp, i32, pd, i64],
        "mrm_cuda_route": [vp, i32, i32, i32, i32, pd, i32, d, pd],
        "mrm_cuda_get_runoff": [vp, i32, i32, pd, i64, i32, i32],
        "mpr_cuda_grid_create": [vp, i32, i32, pi, i32, pi, pi, pi, pi, pi, C.POINTER(vp)],
        "mpr_cuda_grid_destroy": [vp, vp],
        "mpr_cuda_upscale_arithmetic_mean": [vp, vp, d, pd, pd],
        "mpr_cuda_upscale_harmonic_mean": [vp, vp, d, pd, pd],
        "mpr_cuda_upscale_geometric_mean": [vp, vp, d, pd, pd],
        "mpr_cuda_l0_fractional_cover": [vp, vp, pi, i32, pd],
        "mpr_cuda_set_l0": [vp, i32, C.POINTER(MprL0Inputs)],
        "mpr_cuda_set_soildb": [vp, i32, C.POINTER(MprSoilDb)],
        "mpr_cuda_eval": [vp, i32, i32, pd, i32],
        "mhm_cuda_get_param": [vp, i32, i32, i32, pd, i64, i64, i32, i32],
        "mhm_grid_init_lowres_level": [i32, i32, pi, pd, d, d, pi, pi, pi, pi, pi, pd, pi, pi, pi, pi, pi, pi],
        "mhm_cuda_event_record": [vp, i32],
        "mhm_cuda_event_elapsed_ms": [vp, i32, i32, pd],
        "mhm_cuda_synchronize": [vp],
        "mhm_cuda_kernel_stats": [vp, i32, pd, C.POINTER(i64)],
        "mhm_cuda_kernel_stats_reset": [vp, i32],
        "mhm_cuda_measure_dfma_peak": [vp, pd],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    L.mhm_cuda_last_error.restype = C.c_char_p
    L.mhm_cuda_last_error.argtypes = []
    L.mhm_cuda_version.restype = C.c_char_p
    L.mhm_cuda_version.argtypes = []
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise MhmCudaError(load().mhm_cuda_last_error().decode())
