"""Multi-process form of the CPU oracle: the reference's rank-per-chain-group loop
(bipymc/demc.py:39,79-135) with the mpi4py Allgather / Barrier replaced by a shared-memory
population and multiprocessing barriers (mpi4py / mpirun are not in the image).

TEST / BENCH INFRASTRUCTURE ONLY: used by bench.py's `--impl reference` arm and its
cpu_baseline leg to time the reference's algorithm on all host cores.  Each rank owns
np.array_split(range(N), P)[rank], seeds numpy identically (as the reference's tests do,
tests/test_banana.py:17), updates its chains of half a, "all-gathers", updates its chains
of half b -- exactly the reference's schedule, likelihood evaluated twice per step
(samplers.py:330).
"""
import multiprocessing as mp
import time

import numpy as np

from oracle.demc_dream import OracleSampler, _choice_p


class _RankSampler(OracleSampler):
    def attach(self, rank, size, shared, barrier):
        self.rank, self.size, self.shared, self.barrier = rank, size, shared, barrier
        self.ids = np.array_split(np.array(range(self.n_chains)), size)[rank]

    def _allgather(self, cur):
        g = np.frombuffer(self.shared, dtype=np.float64).reshape(self.n_chains, self.dim)
        g[self.ids] = cur[self.ids]
        self.barrier.wait()
        out = g.copy()
        self.barrier.wait()
        return out

    def run_generations(self, gens, **kwargs):
        self._kw = dict(kwargs)
        N = self.n_chains
        flip_prob = np.clip(kwargs.get("flip", 0.5), 0.0, 1.0)
        for k_gen in range(gens):
            fidx, _ = _choice_p([flip_prob, 1 - flip_prob])
            shuffle_idx = np.array(range(N))
            if kwargs.get("shuffle", True):
                np.random.shuffle(shuffle_idx)
            a_ids, b_ids = np.array_split(shuffle_idx, 2)
            if fidx == 0:
                a_ids, b_ids = b_ids, a_ids
            in_a = np.zeros(N, dtype=bool)
            in_a[a_ids] = True
            glob = self._allgather(self.state)              # demc.py:89-94
            self.history[-1] = glob
            cur = glob.copy()
            pool = glob[b_ids].copy()
            for c in self.ids:
                if in_a[c]:
                    cur[c] = self._step(k_gen, c, pool, None, 0)
            glob = self._allgather(cur)                      # demc.py:112-117
            pool = glob[a_ids].copy()
            for c in self.ids:
                if not in_a[c]:
                    cur[c] = self._step(k_gen, c, pool, None, 1)
            cur2 = glob.copy()
            cur2[self.ids] = cur[self.ids]
            self.history.append(cur2)
            self.barrier.wait()                              # demc.py:135


def _worker(rank, size, shared, barrier, spec, gens_warm, gens, out_q):
    from oracle.cases import oracle_target
    fn, kw = oracle_target(spec["target"])
    np.random.seed(spec["seed"])
    s = _RankSampler(fn, spec["theta_0"], n_chains=spec["n_chains"], algo=spec["algo"],
                     varepsilon=spec.get("varepsilon", 1e-6), ln_kwargs=kw, **spec["ctor_kwargs"])
    s.attach(rank, size, shared, barrier)
    import warnings
    warnings.simplefilter("ignore")
    s.run_generations(gens_warm)
    barrier.wait()
    t0 = time.perf_counter()
    s.run_generations(gens)
    barrier.wait()
    dt = time.perf_counter() - t0
    out_q.put((rank, dt, s.local_n_accepted))


def time_port(spec, procs, gens, gens_warm=1):
    """Run `gens` timed generations on `procs` processes; returns (seconds, chain_steps)."""
    ctx = mp.get_context("fork")
    N, d = spec["n_chains"], len(spec["theta_0"])
    shared = ctx.RawArray("d", N * d)
    barrier = ctx.Barrier(procs)
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, procs, shared, barrier, spec, gens_warm, gens, q))
          for r in range(procs)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    return max(r[1] for r in res), N * gens
