#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck; SURVEY.md section 5 "race
detection"): every kernel family runs a few generations at sizes that span several tiles.
  fused_gauss_v4_kernel (TMA-staged gathers, mbarrier + named-barrier hand-over), fused_gauss_v3_kernel,
  split_native / cr_update / flush, gauss_dmma_kernel + propose / accept (d = 1000), fused_small_kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    from bipymc_b200 import DreamMpi, DeMcMpi, targets
    np.random.seed(0)
    for fused in (1, 5):
        s = DreamMpi(targets.Gauss_100D().ln_like, np.zeros(100), n_chains=300, seed=3, varepsilon=0.5,
                     n_cr_gen=1, burnin_gen=100, fused=fused)
        s.run_mcmc(300 * 5)
        h = s._hist.tensor()
        assert h.shape[0] == 5 and bool(np.isfinite(h.cpu().numpy()).all())
        print("gauss100 fused=%d ok, acceptance %.3f" % (fused, s.acceptance_fraction), flush=True)
    s = DreamMpi(targets.Gauss_100D(dim=1000).ln_like, np.zeros(1000), n_chains=160, seed=3, varepsilon=0.5,
                 n_cr_gen=1, burnin_gen=100)
    s.run_mcmc(160 * 3)
    print("gauss1000 ok, acceptance %.3f" % s.acceptance_fraction, flush=True)
    s = DeMcMpi(targets.Gauss_100D().ln_like, np.zeros(100), n_chains=300, seed=3, varepsilon=0.5)
    s.run_mcmc(300 * 5)
    print("gauss100 DE-MC (compile-time one-pair v4 kernel) ok, acceptance %.3f" % s.acceptance_fraction, flush=True)
    s = DeMcMpi(targets.Banana_2D().ln_like, [0.0, 0.0], n_chains=500, seed=3, varepsilon=0.5)
    s.run_mcmc(500 * 6)
    s = DreamMpi(targets.LineFit().ln_like, [-0.8, 4.5, 0.2], n_chains=500, seed=3, varepsilon=1e-2, n_cr_gen=1,
                 burnin_gen=100)
    s.run_mcmc(500 * 6)
    s = DreamMpi(targets.BimodeGauss_2D().ln_like, [0.0, 0.0], n_chains=500, seed=3, varepsilon=0.5, n_cr_gen=1,
                 burnin_gen=100, outlier_gen=2)
    s.run_mcmc(500 * 6)
    print("small-d ok", flush=True)
    print("SANITIZE_CASE_DONE", flush=True)


if __name__ == "__main__":
    main()
