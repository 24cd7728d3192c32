"""Element-wise closeness check used by the parity tests and smoke().  TEST INFRASTRUCTURE (see oracle/__init__.py).

BASELINE.json north_star asks for "Q values, TD errors and classifier probabilities within 1e-4 relative".  TD errors
mix +10000 goal rewards and +1000 option bonuses with ordinary values of size 1-10, so a max-normalised error
(max|a-b| / max|b|) would let a 10 % error on an ordinary element pass.  The bar enforced here is per element:

    |a - b| <= rtol * |b| + rtol * scale(b)

where scale(b) is the RMS of the *typical* elements of b: the non-zero elements no larger than 10x the median
non-zero magnitude (so a few huge rewards do not inflate it).  The second term is what any fp32 computation needs for
elements that are small through cancellation (a sum of F products of size ~scale cannot be known to better than
~eps * scale); it is rtol times the typical magnitude, not the maximum.
"""
import numpy as np


def robust_scale(b):
    m = np.abs(np.asarray(b, dtype=np.float64)).ravel()
    nz = m[m > 0]
    if nz.size == 0:
        return 0.0
    core = nz[nz <= 10.0 * np.median(nz)]
    return float(np.sqrt(np.mean(core * core)))


def mismatch(a, b, rtol=1e-4, scale=None):
    """-> (n_bad, worst ratio err/tol, index of the worst element, scale used)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise AssertionError(f"shape mismatch {a.shape} vs {b.shape}")
    if a.size == 0:
        return 0, 0.0, None, 0.0
    sc = robust_scale(b) if scale is None else float(scale)
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + rtol * sc
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(tol > 0, err / tol, np.where(err > 0, np.inf, 0.0))
    bad = ~(err <= tol)          # NaNs count as mismatches
    i = np.unravel_index(int(np.nanargmax(np.where(np.isnan(ratio), np.inf, ratio))), a.shape)
    return int(bad.sum()), float(ratio[i]), i, sc


def assert_close(a, b, rtol=1e-4, what="", scale=None):
    n_bad, worst, i, sc = mismatch(a, b, rtol, scale)
    if n_bad:
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        raise AssertionError(f"{what or 'values'}: {n_bad} of {a.size} elements outside |a-b| <= {rtol:g}*|b| + "
                             f"{rtol:g}*{sc:.4g}; worst at {i}: got {a[i]!r}, expected {b[i]!r} ({worst:.2f}x the tolerance)")
    return worst
