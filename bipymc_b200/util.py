"""Host-side jitter helpers, same contract as bipymc/util.py:5-28.

They draw from numpy's global legacy stream exactly like the reference, so
``np.random.seed(s)`` followed by a sampler constructor yields the reference's initial
chain states.  Used only for chain initialisation (chain.py:25-29); the per-step jitter
of the hot path is generated inside the CUDA kernels.
"""
import numpy as np


def var_ball(varepsilon, dim):
    """One draw from N(0, diag(varepsilon)); scalar 0. (no RNG consumed) unless all
    variances are > 0."""
    eps = 0.
    v = np.asarray(varepsilon)
    if np.all(v > 0):
        if v.ndim == 0:
            # numpy's multivariate_normal(0, v*I) is sqrt(v) * standard_normal(dim) in
            # draw order (its SVD of v*I is the identity); skip the O(d^3) factorisation.
            eps = np.sqrt(v) * np.random.standard_normal(dim)
        else:
            eps = np.random.multivariate_normal(np.zeros(dim), np.eye(dim) * v, size=1)[0]
    return eps


def var_ball_batch(varepsilon, dim, n):
    """n consecutive var_ball(varepsilon, dim) draws as an (n, dim) array, consuming the
    global stream in the same order as n separate calls (chain after chain)."""
    v = np.asarray(varepsilon)
    if not np.all(v > 0):
        return np.zeros((n, dim))
    if v.ndim == 0:
        return np.sqrt(v) * np.random.standard_normal((n, dim))
    # vector variance: one SVD (numpy sorts singular values, permuting dimensions), then
    # the same affine map numpy applies to every draw
    cov = np.eye(dim) * v
    (_, s, vt) = np.linalg.svd(cov)
    x = np.random.standard_normal((n, dim))
    return np.dot(x, np.sqrt(s)[:, None] * vt)


def var_box(varepsilon, dim):
    eps = 0.
    v = np.asarray(varepsilon)
    if np.all(v > 0):
        eps = np.random.uniform(low=-v * np.ones(dim), high=v * np.ones(dim))
    return eps
