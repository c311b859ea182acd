"""CPU oracle for the serial DeMc sampler (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates bipymc/samplers.py:237-324 (class DeMc: _init_chains :255-259, _mcmc_run :261-308,
_mut_prop_ratio / metropolis_accept :328-336, _super_chain / param_est :311-326) with the
reference's own legacy-np.random call order, so that after np.random.seed(s) it reproduces the
reference bit for bit (pinned by tests/test_oracle_golden.py against tests/golden/ref_serial_*.npz,
which oracle/make_golden.py wrote by running the UNMODIFIED reference).  Only the default
delayed_accept=True schedule is restated: every chain of a sweep proposes from the frozen
previous states and the whole sweep is appended afterwards -- the schedule a batched kernel can
run.  (delayed_accept=False updates chain i before chain i+1 proposes; inherently sequential.)
"""
import numpy as np

from oracle.demc_dream import var_ball, mut_prop_ratio, metropolis_accept


class OracleDeMc(object):
    def __init__(self, ln_like_fn, n_chains=8, ln_kwargs=None):
        assert n_chains >= 4                                         # samplers.py:249
        self.n_chains = n_chains
        ln_kwargs = ln_kwargs or {}
        self._lnl = lambda theta: ln_like_fn(theta, **ln_kwargs)     # samplers.py:43
        self.n_accepted, self.n_rejected = 1, 0                      # samplers.py:30-31
        self.history = []

    @property
    def acceptance_fraction(self):
        return self.n_accepted / (self.n_accepted + self.n_rejected)

    def super_chain(self):
        h = np.array(self.history)                                   # (T, N, d); samplers.py:317-324
        return h.reshape(h.shape[0] * h.shape[1], h.shape[2])

    def param_est(self, n_burn):
        sl = self.super_chain()[n_burn:, :]
        return np.mean(sl, axis=0), np.std(sl, axis=0), sl

    def run_mcmc(self, n, theta_0, varepsilon=1e-6, record=False, **kwargs):
        theta_0 = np.asarray(theta_0, dtype=float)
        N, dim = self.n_chains, len(theta_0)
        gamma = kwargs.get("gamma", 2.38 / np.sqrt(2. * dim))        # samplers.py:264
        assert kwargs.get("delayed_accept", True), "only the delayed-accept schedule is restated"
        # samplers.py:255-259 + chain.py:25-29
        state = np.empty((N, dim))
        for i in range(N):
            state[i] = theta_0 + var_ball(np.asarray(varepsilon * kwargs.get("inflate", 1e1)), dim)
        self.history = [state]
        traces = []
        j = 0
        while j < (n - N):
            cur = self.history[-1]
            new = np.empty_like(cur)
            tr = None
            if record:
                tr = dict(flip=False, shuffle_idx=np.arange(N), pairs=np.zeros((N, 1, 2), dtype=np.int64),
                          gamma_u=np.full(N, np.nan), nrm=np.zeros((N, dim)), accept_u=np.zeros(N),
                          accept=np.zeros(N, dtype=np.int64), prop=np.zeros((N, dim)), lnl_prop=np.zeros(N))
            for i in range(N):
                valid_pool_ids = np.delete(np.array(range(N)), i)
                # np.random.choice(valid_pool_ids, replace=False, size=2) == valid[permutation(N-1)[:2]]
                pick = np.random.permutation(N - 1)[:2]
                mut = valid_pool_ids[pick]
                prop = gamma * (cur[mut[0]] - cur[mut[1]])
                prop += cur[i]
                nrm = var_ball(np.asarray(varepsilon * 1e-3), dim)
                prop += nrm
                lp = self._lnl(prop)
                alpha = mut_prop_ratio(self._lnl(cur[i]), lp)
                acc, u = metropolis_accept(alpha)
                if acc:
                    new[i] = prop
                    self.n_accepted += 1
                else:
                    new[i] = cur[i]
                    self.n_rejected += 1
                if record:
                    tr["pairs"][i, 0] = pick
                    tr["nrm"][i] = nrm
                    tr["accept_u"][i] = u
                    tr["accept"][i] = int(acc)
                    tr["prop"][i] = prop
                    tr["lnl_prop"][i] = lp
                j += 1
            self.history.append(new)
            if record:
                tr["state"] = new.copy()
                traces.append(tr)
        return traces


def replay_serial_generation(X, tr, gamma, lnl_fn):
    """One delayed-accept sweep (samplers.py:271-308) driven by recorded draws `tr`
    (pairs[N,1,2] = positions in np.delete(range(N), i), nrm[N,d], accept_u[N]).
    Returns (new states, accept flags, proposals)."""
    X = np.asarray(X, dtype=float)
    N = X.shape[0]
    new, acc, props = X.copy(), np.zeros(N, dtype=np.int64), np.zeros_like(X)
    for i in range(N):
        valid = np.delete(np.arange(N), i)
        a, b = valid[int(tr["pairs"][i, 0, 0])], valid[int(tr["pairs"][i, 0, 1])]
        prop = gamma * (X[a] - X[b])
        prop += X[i]
        prop += tr["nrm"][i]
        alpha = mut_prop_ratio(lnl_fn(X[i]), lnl_fn(prop))
        if np.isnan(alpha):
            raise ValueError("probabilities contain NaN")
        thr = alpha / (alpha + (1.0 - alpha))                    # numpy choice cdf, p = [alpha, 1 - alpha]
        props[i] = prop
        if tr["accept_u"][i] < thr:
            new[i] = prop
            acc[i] = 1
    return new, acc, props
