#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  The reference
imports mpi4py / h5py / matplotlib / corner at module scope; none are installed, so
oracle/shims/ (a single-rank fake communicator and empty plotting stubs) is put on
sys.path FIRST.  Nothing under /root/reference is copied or modified.

Every case: np.random.seed(seed) -> construct DeMcMpi / DreamMpi -> run_mcmc(n) ->
store the full history (T, N, d), accept counters and (DREAM) the CR-adaptation state.
tests/test_oracle_golden.py replays the same seeds through oracle/demc_dream.py and
demands bit-identical results; the GPU parity tests compare the CUDA path against the
same files.

Usage:  python oracle/make_golden.py                      (rewrites tests/golden/ref_*.npz)
        python oracle/make_golden.py --only=name1,name2   (only those cases)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

import numpy as np  # noqa: E402

from mpi4py import MPI  # noqa: E402  (the shim)
from bipymc.demc import DeMcMpi  # noqa: E402  (the reference)
from bipymc.dream import DreamMpi  # noqa: E402
from bipymc.utils import banana_rv, dblgauss_rv, d100_gauss  # noqa: E402

from bipymc.samplers import DeMc  # noqa: E402

from oracle.cases import ALL_CASES, SERIAL_CASES, linefit_lnprob_ref  # noqa: E402


def target_fn(name):
    if name == "banana":
        return banana_rv.Banana_2D(sigma1=1.0, sigma2=1.0).ln_like, {}
    if name == "dblgauss":
        return dblgauss_rv.BimodeGauss_2D().ln_like, {}
    if name.startswith("gauss"):
        return d100_gauss.Gauss_100D(dim=int(name[5:])).ln_like, {}
    if name == "linefit":
        return linefit_lnprob_ref()
    raise KeyError(name)


def run_case(case):
    fn, ln_kwargs = target_fn(case["target"])
    np.random.seed(case["seed"])
    cls = DreamMpi if case["algo"] == "dream" else DeMcMpi
    s = cls(fn, np.asarray(case["theta_0"], dtype=float), n_chains=case["n_chains"],
            mpi_comm=MPI.COMM_WORLD, ln_kwargs=ln_kwargs, **case["ctor_kwargs"])
    s.run_mcmc(case["n"], **case["run_kwargs"])
    hist = np.array([c.chain for c in s.am_chains])          # (N, T, d)
    hist = np.ascontiguousarray(hist.transpose(1, 0, 2))     # (T, N, d)
    out = dict(history=hist, n_accepted=np.int64(s.n_accepted), n_rejected=np.int64(s.n_rejected),
               acceptance_fraction=np.float64(s.acceptance_fraction))
    if case["algo"] == "dream":
        out.update(p_cr=np.array(s.p_cr), delta_m=np.array(s.delta_m),
                   n_cr_updates=np.array(s.n_cr_updates), CR=np.array(s.CR))
    mean, std, _ = s.param_est(n_burn=0)
    out.update(mean=mean, std=std)
    return out


def run_serial_case(case):
    """Unmodified reference DeMc (samplers.py:237-324), default delayed accept."""
    fn, ln_kwargs = target_fn(case["target"])
    np.random.seed(case["seed"])
    s = DeMc(fn, n_chains=case["n_chains"], ln_kwargs=ln_kwargs)
    s.run_mcmc(case["n"], np.asarray(case["theta_0"], dtype=float), **case["run_kwargs"])
    hist = np.ascontiguousarray(np.array([c.chain for c in s.am_chains]).transpose(1, 0, 2))
    mean, std, sl = s.param_est(n_burn=0)
    return dict(history=hist, n_accepted=np.int64(s.n_accepted), n_rejected=np.int64(s.n_rejected),
                acceptance_fraction=np.float64(s.acceptance_fraction), mean=mean, std=std,
                super_chain_head=sl[:3 * case["n_chains"]])


def main():
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--only=")]
    only = set(only[0]) if only else None
    for name, case in SERIAL_CASES.items():
        if only is not None and name not in only:
            continue
        out = run_serial_case(case)
        path = os.path.join(gdir, "ref_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%-22s T=%d N=%d d=%d acc=%d rej=%d -> %s" % (
            name, out["history"].shape[0], out["history"].shape[1], out["history"].shape[2],
            out["n_accepted"], out["n_rejected"], os.path.relpath(path, ROOT)))
    if "--serial-only" in sys.argv:
        return
    for name, case in ALL_CASES.items():
        if only is not None and name not in only:
            continue
        out = run_case(case)
        path = os.path.join(gdir, "ref_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%-22s T=%d N=%d d=%d acc=%d rej=%d -> %s (%d B)" % (
            name, out["history"].shape[0], out["history"].shape[1], out["history"].shape[2],
            out["n_accepted"], out["n_rejected"], os.path.relpath(path, ROOT), os.path.getsize(path)))


if __name__ == "__main__":
    main()
