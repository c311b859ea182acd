"""Golden-vector case table shared by oracle/make_golden.py (runs the reference) and
tests/ (runs the oracle and the CUDA path).  TEST INFRASTRUCTURE ONLY.

Configurations follow the reference's own tests (tests/test_banana.py:118-127,
tests/test_dblgauss.py:130-140, tests/test_100dgauss.py:100-110) and
examples/ex_para_fit.py:107-112, shortened so the fixtures stay small; DREAM's
n_cr_gen / burnin_gen are shrunk so the fixtures cross every schedule edge
(k % 5 gamma jumps, CR adaptation switching on at len(chain) > n_cr_gen and off at
k >= burnin_gen)."""
import numpy as np


def _case(algo, target, theta_0, n_chains, gens, seed=42, ctor_kwargs=None, run_kwargs=None):
    return dict(algo=algo, target=target, theta_0=list(theta_0), n_chains=n_chains,
                n=n_chains * (gens + 1), gens=gens, seed=seed,
                ctor_kwargs=ctor_kwargs or {}, run_kwargs=run_kwargs or {})


CASES = {
    "banana_demc": _case("demc", "banana", [0.0, 0.0], 20, 60),
    "banana_dream": _case("dream", "banana", [0.0, 0.0], 10, 80,
                          ctor_kwargs=dict(n_cr_gen=5, burnin_gen=40)),
    "banana_dream_odd": _case("dream", "banana", [0.0, 0.0], 11, 30, seed=7,
                              ctor_kwargs=dict(n_cr_gen=3, burnin_gen=20, del_pairs=2, n_cr=4,
                                               gamma_scale=0.9),
                              run_kwargs=dict(flip=0.3, u_epsilon=2e-2, epsilon=1e-9)),
    "dblgauss_demc": _case("demc", "dblgauss", [0.0, 0.0], 20, 40, seed=43,
                           run_kwargs=dict(gamma=0.8, epsilon=1e-6)),
    "dblgauss_dream": _case("dream", "dblgauss", [0.0, 0.0], 10, 60, seed=44,
                            ctor_kwargs=dict(n_cr_gen=10, burnin_gen=200)),
    "gauss100_demc": _case("demc", "gauss100", np.zeros(100), 8, 12, seed=45),
    "gauss100_dream": _case("dream", "gauss100", np.zeros(100), 12, 14, seed=46,
                            ctor_kwargs=dict(n_cr_gen=4, burnin_gen=10)),
    "gauss7_dream_noshuffle": _case("dream", "gauss7", np.zeros(7), 9, 25, seed=47,
                                    ctor_kwargs=dict(n_cr_gen=2, burnin_gen=1000),
                                    run_kwargs=dict(shuffle=False, flip=1.0)),
    "linefit_dream": _case("dream", "linefit", [-0.8, 4.5, 0.2], 12, 50, seed=48,
                           ctor_kwargs=dict(n_cr_gen=10, burnin_gen=30, inflate=1e1)),
}


# Second batch (same recipe): the run-time-general proposal stage at d = 100 / 40 -- DREAM with one and
# with five difference pairs, other n_cr, DE-MC with an explicit gamma, an odd population and two
# k % 10 jump generations; chains start dispersed (varepsilon) so that proposals do get rejected.
# Kept apart from CASES so that the GPU suite runs them last.
EXTRA_CASES = {
    "gauss100_dream_pairs1": _case("dream", "gauss100", np.zeros(100), 10, 12, seed=49,
                                   ctor_kwargs=dict(n_cr_gen=3, burnin_gen=9, del_pairs=1, n_cr=5, varepsilon=0.25)),
    "gauss40_dream_pairs5": _case("dream", "gauss40", np.zeros(40), 16, 10, seed=50,
                                  ctor_kwargs=dict(n_cr_gen=2, burnin_gen=100, del_pairs=5, n_cr=2,
                                                   gamma_scale=1.1, varepsilon=0.25),
                                  run_kwargs=dict(u_epsilon=5e-3)),
    "gauss100_demc_gamma": _case("demc", "gauss100", np.zeros(100), 9, 22, seed=53,
                                 ctor_kwargs=dict(varepsilon=1.0),
                                 run_kwargs=dict(gamma=0.3, epsilon=1e-8, flip=0.7)),
}
# Third batch: a 100-D DREAM population that spans MORE THAN ONE 64-chain tile of the fused kernel
# (136 chains: 68 per half-phase = one full tile + a 4-row partial tile), dispersed start, adaptation
# switching on (len(chain) > 2) and off (k >= 5) inside the fixture.
TILE_CASES = {
    "gauss100_dream_tiles": _case("dream", "gauss100", np.zeros(100), 136, 7, seed=54,
                                  ctor_kwargs=dict(n_cr_gen=2, burnin_gen=5, varepsilon=0.25)),
}
EXTRA_CASES.update(TILE_CASES)
ALL_CASES = dict(CASES, **EXTRA_CASES)


# Serial DeMc (bipymc/samplers.py:237-324, delayed accept): run_mcmc(n, theta_0, **run_kwargs)
SERIAL_CASES = {
    "serial_banana": dict(target="banana", theta_0=[0.0, 0.0], n_chains=12, n=12 * 41, seed=51,
                          run_kwargs=dict(varepsilon=1e-2)),
    "serial_gauss7": dict(target="gauss7", theta_0=list(np.zeros(7)), n_chains=9, n=9 * 21, seed=52,
                          run_kwargs=dict(varepsilon=1e-3, gamma=0.7, inflate=3.0)),
}


def linefit_lnprob_ref():
    """(fn, ln_kwargs) with the signature the reference freezes (samplers.py:43):
    lnprob(theta, x, y, yerr) of examples/ex_para_fit.py:39-55."""
    from oracle.targets import LineFit, linefit_data
    x, y, yerr = linefit_data()

    def lnprob(theta, x, y, yerr):
        return LineFit(x, y, yerr).ln_like(theta)
    return lnprob, dict(x=x, y=y, yerr=yerr)


def oracle_target(name):
    """Restated targets for the oracle / CUDA-parity side (no /root/reference needed)."""
    from oracle import targets
    if name == "banana":
        return targets.Banana2D(sigma1=1.0, sigma2=1.0).ln_like, {}
    if name == "dblgauss":
        return targets.BimodeGauss2D().ln_like, {}
    if name.startswith("gauss"):
        return targets.GaussND(dim=int(name[5:])).ln_like, {}
    if name == "linefit":
        return linefit_lnprob_ref()
    raise KeyError(name)
