"""Parity-harness helpers implementing BASELINE.json's tolerances -- TEST INFRASTRUCTURE ONLY.

Rules:
  * FSQ / LFQ indices: bit-exact.
  * VQ / RVQ indices: may differ from the oracle only on rows where the ORACLE's two
    candidate fp32 distances differ by < 1e-6 relative.
  * quantized outputs, losses, perplexity, EMA buffers: 1e-5 relative (teacher-forced on the
    engine's indices when benign flips exist).
"""
from __future__ import annotations

import numpy as np

INDEX_REL_TOL = 1e-6
VALUE_REL_TOL = 1e-5


def check_indices(idx_test, idx_ref, dist_ref, rel: float = INDEX_REL_TOL):
    """Return (n_flips, n_bad, bad_rows).  A flip on row n is benign iff
    |d[n,i_ref] - d[n,i_test]| / max(|d[n,i_ref]|, |d[n,i_test]|) < rel  using the oracle's
    own fp32 distances `dist_ref` [N,K]."""
    it = np.asarray(idx_test).reshape(-1).astype(np.int64)
    ir = np.asarray(idx_ref).reshape(-1).astype(np.int64)
    assert it.shape == ir.shape, (it.shape, ir.shape)
    K = dist_ref.shape[1]
    assert it.min(initial=0) >= 0 and it.max(initial=0) < K, "index out of range"
    flips = np.nonzero(it != ir)[0]
    if flips.size == 0:
        return 0, 0, flips
    d_ref = dist_ref[flips, ir[flips]].astype(np.float64)
    d_tst = dist_ref[flips, it[flips]].astype(np.float64)
    denom = np.maximum(np.abs(d_ref), np.abs(d_tst))
    with np.errstate(invalid="ignore", divide="ignore"):
        gap = np.where(denom > 0, np.abs(d_ref - d_tst) / denom, 0.0)
    # NaN / inf distances: only an identical bit pattern is acceptable
    same = dist_ref[flips, ir[flips]] == dist_ref[flips, it[flips]]
    bad = ~((gap < rel) | same)
    return int(flips.size), int(bad.sum()), flips[bad]


def rel_err(a, b) -> float:
    """max |a-b| / max(|b|_inf, tiny): a tensor-level relative error (the scale of the tensor,
    not of each element -- elements that cancel to ~0 are judged against the tensor's scale)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {a.shape} vs {b.shape}")
    if a.size == 0:
        return 0.0
    if not np.array_equal(np.isfinite(a), np.isfinite(b)):
        return float("inf")
    m = np.isfinite(b)
    if not m.any():
        return 0.0
    scale = max(float(np.max(np.abs(b[m]))), 1e-30)
    return float(np.max(np.abs(a[m] - b[m])) / scale)


def rel_err_rows(a, b) -> float:
    """Row-wise variant: each row (last axis) judged against its own max magnitude.  Used for
    codebooks whose rows span 10 orders of magnitude in the fresh-init regime."""
    a = np.asarray(a, np.float64).reshape(-1, np.asarray(a).shape[-1])
    b = np.asarray(b, np.float64).reshape(a.shape)
    scale = np.maximum(np.max(np.abs(b), axis=1, keepdims=True), 1e-30)
    return float(np.max(np.abs(a - b) / scale))


def assert_close(a, b, tol: float = VALUE_REL_TOL, what: str = "", rows: bool = False):
    e = rel_err_rows(a, b) if rows else rel_err(a, b)
    if not e <= tol:
        raise AssertionError(f"{what}: relative error {e:.3e} > {tol:.1e}")
    return e
