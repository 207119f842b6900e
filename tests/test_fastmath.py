"""Accuracy of the fast cell kernel's log / exp / pow (mhm_b200/csrc/fastmath.cuh) -- their host
instantiations, exported by the library for this test -- against glibc on ~1e6 samples."""
import ctypes as C

import numpy as np

from mhm_b200 import _lib


def ulp_err(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_fast_log_exp_pow_accuracy():
    L = _lib.load()
    for f, n in (("mhm_host_fast_log", 1), ("mhm_host_fast_exp", 1), ("mhm_host_fast_pow", 2)):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double] * n
    rng = np.random.default_rng(0)
    N = 200000
    # log over the kernel's argument range: ratios sm/sat in (1e-18, 1], storages up to 1e5
    x = np.concatenate([10.0 ** rng.uniform(-18, 5, N), rng.uniform(0.5, 2.0, N), 1.0 + rng.normal(0, 1e-6, 1000)])
    got = np.array([L.mhm_host_fast_log(v) for v in x])
    ref = np.log(x)
    e = ulp_err(got[ref != 0], ref[ref != 0])
    assert e.max() <= 1.0, e.max()
    # exp over the exponent range of b*log(ratio) and (1+alpha)*log(S)
    t = np.concatenate([rng.uniform(-300, 30, N), rng.uniform(-1, 1, N)])
    got = np.array([L.mhm_host_fast_exp(v) for v in t])
    e = ulp_err(got, np.exp(t))
    assert e.max() <= 1.0, e.max()
    # pow: exp(y*log x) as the reference writes it; the rounding of log and of the product are
    # amplified by |y log x|: relative error <= 2.3e-16 * (1 + |y log x|)
    xs = np.concatenate([rng.uniform(1e-6, 1.0, N), rng.uniform(1e-12, 5e3, N)])
    ys = np.concatenate([rng.uniform(1.5, 6.0, N), rng.uniform(1.05, 1.6, N)])
    got = np.array([L.mhm_host_fast_pow(a, b) for a, b in zip(xs, ys)])
    ref = np.power(xs, ys)
    rel = np.abs(got - ref) / ref
    bound = 2.3e-16 * (1.0 + np.abs(ys * np.log(xs)))
    assert (rel <= bound).all(), (rel / bound).max()
    assert L.mhm_host_fast_pow(1.0, 3.7) == 1.0 and L.mhm_host_fast_exp(0.0) == 1.0


def test_fast_pow23_accuracy():
    """canopy evaporation's x**(2/3) via the Newton-refined inverse cube root: <= 2 ulp"""
    L = _lib.load()
    L.mhm_host_fast_pow23.restype = C.c_double
    L.mhm_host_fast_pow23.argtypes = [C.c_double]
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-29, 0, 200000), rng.uniform(0.0, 1.0, 100000) + 1e-300,
                        10.0 ** rng.uniform(-300, -30, 1000)])
    got = np.array([L.mhm_host_fast_pow23(v) for v in x])
    ref = np.array([float(np.longdouble(v) ** (np.longdouble(2) / 3)) for v in x])
    e = ulp_err(got, ref)
    assert e.max() <= 2.0, e.max()
    assert L.mhm_host_fast_pow23(1.0) == 1.0


def test_table_driven_log_exp_pow_accuracy():
    """the 128-entry table versions the fast kernel uses for x**y: log absolute error within
    1 ulp(|log x|) + 6e-17, exp <= 1 ulp, pow relative error <= 3e-16 * (1 + |y log x|)"""
    L = _lib.load()
    for f, n in (("mhm_host_tab_log", 1), ("mhm_host_tab_exp", 1), ("mhm_host_tab_pow", 2)):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double] * n
    rng = np.random.default_rng(2)
    N = 200000
    x = np.concatenate([10.0 ** rng.uniform(-18, 5, N), rng.uniform(0.5, 2.0, N), 1.0 + rng.normal(0, 1e-6, 1000),
                        10.0 ** rng.uniform(-300, 300, 20000)])
    got = np.array([L.mhm_host_tab_log(v) for v in x])
    ref = np.log(x)
    assert (np.abs(got - ref) <= np.spacing(np.abs(ref)) + 6e-17).all()
    t = np.concatenate([rng.uniform(-690, 690, N), rng.uniform(-1, 1, N)])
    got = np.array([L.mhm_host_tab_exp(v) for v in t])
    assert ulp_err(got, np.exp(t)).max() <= 1.0
    xs = np.concatenate([rng.uniform(1e-6, 1.0, N), rng.uniform(1e-12, 5e3, N), 1.0 - 10.0 ** rng.uniform(-16, -1, 20000)])
    ys = np.concatenate([rng.uniform(1.5, 6.0, N), rng.uniform(1.05, 1.6, N), rng.uniform(1.0, 6.0, 20000)])
    got = np.array([L.mhm_host_tab_pow(a, b) for a, b in zip(xs, ys)])
    ref = np.power(xs, ys)
    rel = np.abs(got - ref) / ref
    assert (rel <= 3.0e-16 * (1.0 + np.abs(ys * np.log(xs)))).all()
    assert L.mhm_host_tab_pow(1.0, 3.7) == 1.0 and L.mhm_host_tab_exp(0.0) == 1.0
