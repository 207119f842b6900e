"""mhm_b200 -- B200 (sm_100a) implementation of mHM's L1 hot path behind a C ABI.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/mhm_cuda.h),
interface.py (host-side mirror of the reference's run interface), synth.py (synthetic
domains of the BASELINE shapes) and fortran/ (the ISO_C_BINDING module the reference's
driver would use).  Importing the package does not need a GPU; creating a Context does.
"""
from ._lib import MhmCudaError, load  # noqa: F401
from .interface import Context, Domain, routing_order, time_indices  # noqa: F401

__all__ = ["Context", "Domain", "MhmCudaError", "load", "routing_order", "time_indices"]
