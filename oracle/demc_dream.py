"""CPU oracle for the DE-MC / DREAM per-generation update (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (bipymc_b200/) never does.

This is a restatement, in plain numpy with per-chain Python loops, of

  * DeMcMpi._mcmc_run            bipymc/demc.py:63-151   (generation loop, a/b pools)
  * DeMcMpi._update_chain_pool   bipymc/demc.py:153-196  (DE-MC step)
  * DreamMpi._update_chain_pool  bipymc/dream.py:32-107  (DREAM step)
  * DreamMpi._init_cr / _update_cr_ratios  bipymc/dream.py:109-140
  * DeMc._mut_prop_ratio / metropolis_accept  bipymc/samplers.py:328-336
  * util.var_ball / var_box      bipymc/util.py:5-28
  * McmcChain.__init__           bipymc/chain.py:13-29   (initial jitter)

for a single rank (comm.size == 1, the only bit-reproducible configuration of the
reference).  It draws from the global legacy ``np.random`` stream
with the same calls in the same order as the reference, so that after
``np.random.seed(s)`` it reproduces the reference's chains BIT FOR BIT.  That claim is
pinned by tests/test_oracle_golden.py against tests/golden/ref_*.npz, which were
written by oracle/make_golden.py running the UNMODIFIED reference from /root/reference.

While stepping it can record every random draw of every chain-step into flat replay
buffers (``record=True``); the CUDA kernels consume those buffers in RNG-replay mode
and must then reproduce accept decisions exactly and states to 1e-12 relative.

Two deliberate shortcuts, both verified bit-exact by the golden pin:
  * ``np.random.choice([x, y], p=[p0, p1])`` is restated as one ``random_sample()``
    pushed through numpy's own cdf/searchsorted rule, so the uniform can be recorded;
  * ``var_ball(v, d)`` with scalar v is restated as ``sqrt(v) * standard_normal(d)``
    (numpy's multivariate_normal does an SVD of v*I first; same numbers, far cheaper --
    this makes the CPU baseline FASTER than the reference, i.e. conservative).
"""
import numpy as np


def _choice_p(p):
    """One draw of np.random.choice(range(len(p)), p=p): returns (index, uniform).

    numpy legacy RandomState.choice: cdf = p.cumsum(); cdf /= cdf[-1];
    u = random_sample(); idx = cdf.searchsorted(u, side='right')."""
    cdf = np.asarray(p, dtype=float).cumsum()
    cdf /= cdf[-1]
    u = np.random.random_sample()
    return int(cdf.searchsorted(u, side="right")), u


def var_ball(varepsilon, dim):
    """util.py:5-16.  Scalar variance only on the hot path (epsilon**2)."""
    v = np.asarray(varepsilon)
    if np.all(v > 0):
        if v.ndim == 0:
            return np.sqrt(v) * np.random.standard_normal(dim)
        return np.random.multivariate_normal(np.zeros(dim), np.eye(dim) * v, size=1)[0]
    return 0.0


def var_box(varepsilon, dim):
    """util.py:18-28."""
    v = np.asarray(varepsilon)
    if np.all(v > 0):
        return np.random.uniform(low=-v * np.ones(dim), high=v * np.ones(dim))
    return 0.0


def mut_prop_ratio(lnl_cur, lnl_prop):
    """samplers.py:328-332 with the two likelihood values already evaluated."""
    with np.errstate(over="ignore", invalid="ignore"):
        alpha = np.min((1.0, np.exp(lnl_prop - lnl_cur)))
    return np.clip(alpha, 0.0, 1.0)


def metropolis_accept(alpha):
    """samplers.py:334-336: np.random.choice([True, False], p=[alpha, 1-alpha]).
    Returns (accept, uniform).  NaN alpha raises ValueError like numpy does."""
    if np.isnan(alpha):
        raise ValueError("probabilities contain NaN")
    idx, u = _choice_p([alpha, 1.0 - alpha])
    return idx == 0, u


class OracleSampler(object):
    """Single-rank DeMcMpi / DreamMpi restatement.  algo in {"demc", "dream"}."""

    def __init__(self, ln_like_fn, theta_0, n_chains=8, algo="dream", varepsilon=1e-6,
                 ln_kwargs=None, **kwargs):
        assert n_chains >= 4                                  # samplers.py:249
        assert algo in ("demc", "dream")
        self.algo = algo
        self.n_chains = n_chains
        ln_kwargs = ln_kwargs or {}
        self._lnl = lambda theta: ln_like_fn(theta, **ln_kwargs)   # samplers.py:43
        theta_0 = np.asarray(theta_0, dtype=float).flatten()
        self.dim = len(theta_0)
        # dream.py:20-27
        self.gamma_scale = kwargs.get("gamma_scale", 1.0)
        self.del_pairs = kwargs.get("del_pairs", 3)
        self.burnin_gen = kwargs.get("burnin_gen", 300)
        self.p_cr_update_gen = kwargs.get("n_cr_gen", 50)
        self.n_cr = kwargs.get("n_cr", 3)
        # chain.py:25-29, one McmcChain per chain in ascending id order
        varepsilon = np.asarray(varepsilon)
        state0 = np.empty((n_chains, self.dim))
        for c in range(n_chains):
            state0[c] = theta_0 + var_ball(varepsilon, self.dim)
        self.history = [state0]            # list of (N, d): row t of every chain
        self.local_n_accepted, self.local_n_rejected = 0, 1   # demc.py:19-20
        self.n_accepted, self.n_rejected = 1, 0               # samplers.py:30-31
        if algo == "dream":
            self._init_cr()

    # ---- dream.py:109-117
    def _init_cr(self):
        self.CR = (np.array(range(self.n_cr)) + 1) / self.n_cr
        self.p_cr = np.ones(self.n_cr) / self.n_cr
        self.n_cr_updates = np.zeros(self.n_cr)
        self.p_cr_update = np.zeros(self.n_cr)
        self.delta_m = np.zeros(self.n_cr)

    @property
    def state(self):
        return self.history[-1]

    def chain(self, c):
        """(T, d) history of chain c (McmcChain.chain)."""
        return np.array([h[c] for h in self.history])

    @property
    def hist_array(self):
        return np.array(self.history)          # (T, N, d)

    @property
    def acceptance_fraction(self):
        return self.n_accepted / (self.n_accepted + self.n_rejected)

    def super_chain(self):
        """demc.py:260-268: row t*N + i = chain i at time t."""
        h = self.hist_array
        return h.reshape(h.shape[0] * h.shape[1], h.shape[2])

    def param_est(self, n_burn):
        sl = self.super_chain()[n_burn:, :]
        return np.mean(sl, axis=0), np.std(sl, axis=0), sl

    # ---- demc.py:63-151
    def run_mcmc(self, n, record=False, max_gen=None, **kwargs):
        N, d = self.n_chains, self.dim
        self.local_n_accepted, self.local_n_rejected = 0, 1
        flip_prob = np.clip(kwargs.get("flip", 0.5), 0.0, 1.0)
        shuffle = kwargs.get("shuffle", True)
        traces = []
        j, k_gen = 0, 0
        while j < int((n - N) / 1):
            if max_gen is not None and k_gen >= max_gen:
                break
            fidx, flip_u = _choice_p([flip_prob, 1 - flip_prob])
            flip_bool = (fidx == 0)
            shuffle_idx = np.array(range(N))
            if shuffle:
                np.random.shuffle(shuffle_idx)
            a_ids, b_ids = np.array_split(shuffle_idx, 2)
            if flip_bool:
                a_ids, b_ids = b_ids, a_ids
            tr = None
            if record:
                tr = self._new_trace(k_gen, flip_bool, flip_u, shuffle_idx, a_ids, b_ids)
            cur = self.state.copy()          # becomes row t+1 as chains are updated
            # phase a: pool is the frozen b half (demc.py:95-109)
            pool = self.state[b_ids].copy()
            in_a = np.zeros(N, dtype=bool)
            in_a[a_ids] = True
            for c in range(N):
                if not in_a[c]:
                    continue
                j += 1
                cur[c] = self._step(k_gen, c, pool, tr, 0)
            # phase b: pool is the UPDATED a half (demc.py:112-132)
            pool = cur[a_ids].copy()
            for c in range(N):
                if in_a[c]:
                    continue
                j += 1
                cur[c] = self._step(k_gen, c, pool, tr, 1)
            self.history.append(cur)
            k_gen += 1
            if record:
                tr["state"] = cur.copy()
                if self.algo == "dream":
                    tr["p_cr"] = self.p_cr.copy()
                    tr["delta_m"] = self.delta_m.copy()
                    tr["n_cr_updates"] = self.n_cr_updates.copy()
                traces.append(tr)
        self.n_accepted = self.local_n_accepted          # demc.py:143-150, size == 1
        self.n_rejected = self.local_n_rejected
        return traces

    def _new_trace(self, k, flip_bool, flip_u, shuffle_idx, a_ids, b_ids):
        N, d = self.n_chains, self.dim
        npair = self.del_pairs if self.algo == "dream" else 1
        tr = dict(k=k, flip=bool(flip_bool), flip_u=flip_u, shuffle_idx=shuffle_idx.copy(),
                  a_ids=a_ids.copy(), b_ids=b_ids.copy(),
                  phase=np.zeros(N, dtype=np.int32),
                  pairs=np.zeros((N, npair, 2), dtype=np.int32),
                  gamma_u=np.full(N, np.nan), nrm=np.zeros((N, d)), accept_u=np.zeros(N),
                  accept=np.zeros(N, dtype=np.int32), alpha=np.zeros(N),
                  lnl_cur=np.zeros(N), lnl_prop=np.zeros(N), prop=np.zeros((N, d)))
        if self.algo == "dream":
            tr.update(cr_idx=np.zeros(N, dtype=np.int32), z=np.zeros((N, d)),
                      fallback_dim=np.full(N, -1, dtype=np.int32), e=np.zeros((N, d)),
                      p_cr_in=self.p_cr.copy())
        return tr

    def _step(self, k, c, pool, tr, phase):
        if self.algo == "dream":
            return self._step_dream(k, c, pool, tr, phase)
        return self._step_demc(k, c, pool, tr, phase)

    # ---- demc.py:153-196
    def _step_demc(self, k, c, pool, tr, phase):
        kw = self._kw
        epsilon = kw.get("epsilon", 1e-15)
        gamma_base = kw.get("gamma", 2.38 / np.sqrt(2. * self.dim))
        cur = self.state[c]
        ids = np.random.choice(np.array(range(len(pool))), replace=False, size=2)
        gamma_u = np.nan
        if k % 10 == 0:
            gi, gamma_u = _choice_p([0.1, 0.9])
            gamma = [gamma_base, 1.0][gi]
        else:
            gamma = gamma_base
        prop = gamma * (pool[ids[0]] - pool[ids[1]])
        prop += cur
        nrm = var_ball(epsilon ** 2.0, self.dim)
        prop += nrm
        return self._finish(c, cur, prop, tr, phase, dict(pairs=ids.reshape(1, 2), gamma_u=gamma_u, nrm=nrm))

    # ---- dream.py:32-107
    def _step_dream(self, k, c, pool, tr, phase):
        kw = self._kw
        d = self.dim
        epsilon = kw.get("epsilon", 1e-12)
        u_epsilon = kw.get("u_epsilon", 1e-2)
        cur = self.state[c]
        valid = np.array(range(len(pool)))
        cr_idx, _ = _choice_p(self.p_cr)                       # dream.py:51
        cr = self.CR[cr_idx]
        z = np.random.uniform(0, 1, size=d)                    # :52
        mask = (z <= cr)
        fallback = -1
        if np.count_nonzero(mask) == 0:                        # :55-57
            fallback = int(np.random.choice(range(len(mask))))
            mask[fallback] = True
        d_prime = np.count_nonzero(mask)
        gamma_base = self.gamma_scale * 2.38 / np.sqrt(2. * self.del_pairs * d_prime)   # :61
        np.random.choice(valid, replace=True, size=(2, self.del_pairs))                 # :62 dead draw
        pairs = np.zeros((self.del_pairs, 2), dtype=np.int64)
        A = np.zeros((self.del_pairs, d))
        B = np.zeros((self.del_pairs, d))
        for p in range(self.del_pairs):                        # :65-68
            ids = np.random.choice(valid, replace=False, size=(2,))
            pairs[p] = ids
            A[p] = pool[ids[0]]
            B[p] = pool[ids[1]]
        update_dims = np.zeros(d)
        update_dims[mask] = 1.0
        gamma_u = np.nan
        if k % 5 == 0:                                         # :77-80
            gi, gamma_u = _choice_p([0.20, 0.80])
            gamma = [gamma_base, 1.0][gi]
        else:
            gamma = gamma_base
        eps_u = var_box(u_epsilon, d)                          # :83
        eps_n = var_ball(epsilon ** 2.0, d)                    # :84
        prop = ((np.ones(d) + eps_u) * gamma * np.sum(A - B, axis=0) + eps_n) * update_dims   # :85-86
        prop += cur                                            # :89
        if self.burnin_gen > k:                                # :92-93
            self._update_cr_ratios(c, cur, prop, cr_idx)
        return self._finish(c, cur, prop, tr, phase,
                            dict(pairs=pairs, gamma_u=gamma_u, nrm=eps_n, cr_idx=cr_idx, z=z,
                                 fallback_dim=fallback, e=eps_u))

    # ---- dream.py:119-140
    def _update_cr_ratios(self, c, cur, prop, cr_idx):
        n_gen = len(self.history)
        if n_gen > self.p_cr_update_gen:
            self.n_cr_updates[cr_idx] += 1.0
            std_devs = np.std(self.chain(c), axis=0)
            std_devs[std_devs == 0] = 1e-12
            self.delta_m[cr_idx] += np.sum(((cur - prop) ** 2.0 / std_devs ** 2.0))
            if np.count_nonzero(self.n_cr_updates) == self.n_cr:
                for m in range(self.n_cr):
                    self.p_cr_update[m] = (self.delta_m[m] / self.n_cr_updates[m])
                self.p_cr = self.p_cr_update
            self.p_cr /= np.sum(self.p_cr)

    def _finish(self, c, cur, prop, tr, phase, rec):
        lnl_prop = self._lnl(prop)           # samplers.py:330 evaluates proposal first,
        lnl_cur = self._lnl(cur)             # then the current state (never cached)
        alpha = mut_prop_ratio(lnl_cur, lnl_prop)
        accept, u = metropolis_accept(alpha)
        if accept:
            new_state = prop
            self.local_n_accepted += 1
        else:
            new_state = cur
            self.local_n_rejected += 1
        if tr is not None:
            tr["phase"][c] = phase
            for key, val in rec.items():
                tr[key][c] = val
            tr["accept_u"][c] = u
            tr["accept"][c] = int(accept)
            tr["alpha"][c] = alpha
            tr["lnl_cur"][c] = lnl_cur
            tr["lnl_prop"][c] = lnl_prop
            tr["prop"][c] = prop
        return np.array(new_state, dtype=float)

    _kw = {}

    def run(self, n, **kwargs):
        """Convenience: run_mcmc with the kwargs also visible to the step functions
        (the reference forwards run_mcmc's **kwargs to _update_chain_pool, demc.py:109)."""
        record = kwargs.pop("record", False)
        max_gen = kwargs.pop("max_gen", None)
        self._kw = dict(kwargs)
        try:
            return self.run_mcmc(n, record=record, max_gen=max_gen, **kwargs)
        finally:
            self._kw = {}
