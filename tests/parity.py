"""Comparison helpers for the parity tests.

Tolerance (BASELINE.json north_star): gridded fluxes and states <= 1e-9 relative per
timestep; gauge discharge <= 1e-8 relative.  A pure relative test is meaningless for values
that are differences of O(1) quantities (e.g. aET = pet - aet_canopy - ...), so the check is
|a - b| <= RTOL * max(|a|, |b|) + ATOL with ATOL = 1e-12 (mm or m3/s), i.e. eight orders
below the smallest physically meaningful flux.  The strict relative violations are counted
and reported, never hidden.
"""
import numpy as np

RTOL = 1e-9
ATOL = 1e-12
RTOL_Q = 1e-8


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(a), np.abs(b))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(den > 0, np.abs(a - b) / den, 0.0)
    r = np.where(np.isnan(a) & np.isnan(b), 0.0, r)
    return r


def assert_close(got, ref, what, rtol=RTOL, atol=ATOL):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    nan_mismatch = np.isnan(got) != np.isnan(ref)
    assert not nan_mismatch.any(), "%s: NaN pattern differs at %d places" % (what, nan_mismatch.sum())
    bad = np.abs(got - ref) > rtol * np.maximum(np.abs(got), np.abs(ref)) + atol
    bad &= ~np.isnan(ref)
    if bad.any():
        i = np.argmax(np.where(bad, np.abs(got - ref), 0.0))
        raise AssertionError("%s: %d of %d values differ; worst got=%r ref=%r" % (
            what, bad.sum(), bad.size, got.flat[i], ref.flat[i]))
    return float(rel_err(got, ref).max()) if got.size else 0.0


def assert_bit_exact(got, ref, what):
    got, ref = np.ascontiguousarray(got), np.ascontiguousarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    same = got.view(np.uint64) == ref.view(np.uint64) if got.dtype == np.float64 else got == ref
    if not same.all():
        i = int(np.argmin(same))
        raise AssertionError("%s: %d of %d values not bit-identical; first got=%r ref=%r" % (
            what, (~same).sum(), same.size, got.flat[i], ref.flat[i]))
