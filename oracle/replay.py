"""Replay-driven batched form of the oracle (TEST INFRASTRUCTURE ONLY).

Same arithmetic as oracle/demc_dream.py (which is pinned bit-for-bit to the reference),
but every random draw is READ from a trace dict instead of drawn from np.random, and the
chains of one half-phase are processed as numpy vectors.  All per-dimension operations
are elementwise float64 in the reference's order, so results are bit-identical to the
scalar oracle given the same draws (tests/test_oracle_golden.py checks that).

Used to (a) check the CUDA native-RNG mode: the device dumps the draws its Philox stream
will use (bpm_dump_draws), this module replays them on the CPU; (b) run parity at sizes
where the per-chain Python loop of the scalar oracle would take minutes.

Follows bipymc/demc.py:79-135,161-196 and bipymc/dream.py:40-107,119-140.
"""
import numpy as np


def _cdf_pick(p, u):
    cdf = np.asarray(p, dtype=float).cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(u, side="right")


def split_ids(shuffle_idx, flip):
    a_ids, b_ids = np.array_split(np.asarray(shuffle_idx), 2)     # demc.py:95-97
    if flip:
        a_ids, b_ids = b_ids, a_ids                                # demc.py:98-100
    return a_ids, b_ids


class CrState(object):
    """dream.py:109-117"""
    def __init__(self, n_cr):
        self.n_cr = n_cr
        self.CR = (np.array(range(n_cr)) + 1) / n_cr
        self.p_cr = np.ones(n_cr) / n_cr
        self.n_cr_updates = np.zeros(n_cr)
        self.delta_m = np.zeros(n_cr)

    def batch_update(self, cr_idx, delta):
        """Generation-boundary form of dream.py:119-140: add every chain's jump statistic,
        then (if every CR value has been used) p_cr = delta_m / n_cr_updates, normalised."""
        if len(cr_idx) == 0:
            return
        for m in range(self.n_cr):
            sel = (cr_idx == m)
            self.n_cr_updates[m] += np.count_nonzero(sel)
            self.delta_m[m] += np.sum(delta[sel])
        if np.count_nonzero(self.n_cr_updates) == self.n_cr:
            self.p_cr = self.delta_m / self.n_cr_updates
        self.p_cr = self.p_cr / np.sum(self.p_cr)


def replay_generation(X, tr, cfg, lnl_batch, k, hist_var=None, cr=None, lnl_cache=None):
    """One generation from recorded draws.

    X          (N, d) state before the generation
    tr         trace dict (keys as written by OracleSampler(record=True) or
               bipymc_b200's dump of the native stream): flip, shuffle_idx, pairs,
               gamma_u, nrm, accept_u [, cr_idx, z, fallback_dim, e]
    cfg        dict: algo, del_pairs, n_cr, gamma_scale, gamma (DE-MC, may be None),
               burnin_gen, n_cr_gen
    lnl_batch  callable (n, d) -> (n,)
    hist_var   (N, d) population variance of every chain's history (np.std(...)**2 is
               what the reference uses) or None when CR adaptation is inactive
    cr         CrState (DREAM), updated in place at the end of the generation
    Returns dict(state, accept, lnl_prop, prop, alpha).
    """
    N, d = X.shape
    dream = cfg["algo"] == "dream"
    a_ids, b_ids = split_ids(tr["shuffle_idx"], tr["flip"])
    cur = X.copy()
    out = dict(accept=np.zeros(N, dtype=np.int32), lnl_prop=np.zeros(N), prop=np.zeros((N, d)),
               alpha=np.zeros(N))
    adapt_idx, adapt_delta = [], []
    for phase, (self_ids, pool_ids) in enumerate(((a_ids, b_ids), (b_ids, a_ids))):
        c = np.asarray(self_ids)
        pool = cur[pool_ids]                     # phase 1 sees the UPDATED a half
        xc = cur[c]
        pairs = np.asarray(tr["pairs"])[c]       # (n, npair, 2) pool-local
        nrm = np.asarray(tr["nrm"])[c]
        gu = np.asarray(tr["gamma_u"])[c]
        if dream:
            cr_idx = np.asarray(tr["cr_idx"])[c]
            crv = ((np.array(range(cfg["n_cr"])) + 1) / cfg["n_cr"])[cr_idx]
            z = np.asarray(tr["z"])[c]
            mask = z <= crv[:, None]
            empty = np.count_nonzero(mask, axis=1) == 0
            if np.any(empty):
                fb = np.asarray(tr["fallback_dim"])[c]
                mask[np.where(empty)[0], fb[empty]] = True
            d_prime = np.count_nonzero(mask, axis=1)
            gamma_base = cfg["gamma_scale"] * 2.38 / np.sqrt(2. * cfg["del_pairs"] * d_prime)
            if k % 5 == 0:
                gamma = np.where(gu < _cdf_pick_thr([0.20, 0.80]), gamma_base, 1.0)
            else:
                gamma = gamma_base
            S = pool[pairs[:, 0, 0]] - pool[pairs[:, 0, 1]]
            for p in range(1, cfg["del_pairs"]):
                S = S + (pool[pairs[:, p, 0]] - pool[pairs[:, p, 1]])
            e = np.asarray(tr["e"])[c]
            prop = ((np.ones(d) + e) * gamma[:, None] * S + nrm) * mask.astype(float)
            prop = prop + xc
            if cfg["burnin_gen"] > k and hist_var is not None:
                var = hist_var[c].copy()
                var[var == 0] = 1e-12 ** 2.0
                adapt_idx.append(cr_idx)
                adapt_delta.append(np.sum((xc - prop) ** 2.0 / var, axis=1))
        else:
            gamma_base = cfg.get("gamma") or 2.38 / np.sqrt(2. * d)
            if k % 10 == 0:
                gamma = np.where(gu < _cdf_pick_thr([0.1, 0.9]), gamma_base, 1.0)[:, None]
            else:
                gamma = gamma_base
            prop = gamma * (pool[pairs[:, 0, 0]] - pool[pairs[:, 0, 1]])
            prop = prop + xc
            prop = prop + nrm
        lp = np.asarray(lnl_batch(prop), dtype=float)
        lc = np.asarray(lnl_batch(xc), dtype=float) if lnl_cache is None else lnl_cache[c]
        with np.errstate(over="ignore", invalid="ignore"):
            alpha = np.minimum(1.0, np.exp(lp - lc))
        alpha = np.clip(alpha, 0.0, 1.0)
        if np.any(np.isnan(alpha)):
            raise ValueError("probabilities contain NaN")
        thr = alpha / (alpha + (1.0 - alpha))
        acc = np.asarray(tr["accept_u"])[c] < thr
        cur[c[acc]] = prop[acc]
        if lnl_cache is not None:
            lnl_cache[c[acc]] = lp[acc]
        out["accept"][c] = acc
        out["lnl_prop"][c] = lp
        out["prop"][c] = prop
        out["alpha"][c] = alpha
    if dream and cr is not None and adapt_idx:
        cr.batch_update(np.concatenate(adapt_idx), np.concatenate(adapt_delta))
    out["state"] = cur
    return out


def _cdf_pick_thr(p):
    cdf = np.asarray(p, dtype=float).cumsum()
    cdf /= cdf[-1]
    return cdf[0]


def scalar_batch(fn):
    """Wrap a scalar ln_like(theta) as a batch callable."""
    def f(rows):
        return np.array([float(fn(r)) for r in np.asarray(rows)])
    return f
