"""CPU oracle for the burn-in diagnostics (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

The reference has neither an outlier-chain reset nor an R-hat; BASELINE.json's north_star asks
for both, "taken from Vrugt et al. 2008/2009" (the papers the reference cites, readme.md:41-43).
Parity for these two is therefore pinned to the published definitions restated here in numpy
(parity UNPINNED by the reference itself: it holds no vectors for them):

  * IQR outlier reset (Vrugt et al. 2009, section 3.3): Omega_c = mean log-density of chain c
    over a trailing window; outliers are chains with Omega_c < Q1 - 2 (Q3 - Q1); an outlier
    takes the current state of the best chain.
  * Gelman-Rubin R-hat (Gelman & Rubin 1992) per dimension.
"""
import numpy as np


def iqr_outliers(omega):
    """(mask of outlier chains, threshold, best chain)."""
    omega = np.asarray(omega, dtype=float)
    q1, q3 = np.percentile(omega, [25.0, 75.0])
    thr = q1 - 2.0 * (q3 - q1)
    best = int(np.argmax(omega))
    out = omega < thr
    out[best] = False
    return out, thr, best


def outlier_reset(X, lnl, omega):
    """Returns (X', lnl', mask): outlier rows replaced by the best chain's row."""
    X = np.array(X, dtype=float, copy=True)
    lnl = np.array(lnl, dtype=float, copy=True)
    out, thr, best = iqr_outliers(omega)
    X[out] = X[best]
    lnl[out] = lnl[best]
    return X, lnl, out


def rhat(hist):
    """hist: (T, N, d) rows of every chain.  R-hat per dimension."""
    hist = np.asarray(hist, dtype=float)
    T, N, d = hist.shape
    cm = hist.mean(axis=0)                        # (N, d)
    W = hist.var(axis=0, ddof=1).mean(axis=0)
    B_over_T = cm.var(axis=0, ddof=1)
    return np.sqrt(((T - 1.0) / T * W + B_over_T) / W)
