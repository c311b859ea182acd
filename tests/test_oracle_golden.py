"""The oracle pin: oracle/demc_dream.py must reproduce, BIT FOR BIT, the chains the
unmodified reference produced for the same seeds (tests/golden/ref_*.npz, written by
oracle/make_golden.py from /root/reference).  Also pins the replay-driven batched form
(oracle/replay.py) to the scalar oracle."""
import os
import warnings

import numpy as np
import pytest

from oracle.cases import ALL_CASES, oracle_target
from oracle.demc_dream import OracleSampler
from oracle import replay as orp

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_oracle(name, record=False):
    case = ALL_CASES[name]
    fn, kw = oracle_target(case["target"])
    np.random.seed(case["seed"])
    s = OracleSampler(fn, case["theta_0"], n_chains=case["n_chains"], algo=case["algo"],
                      ln_kwargs=kw, **case["ctor_kwargs"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tr = s.run(case["n"], record=record, **case["run_kwargs"])
    return s, tr


@pytest.mark.parametrize("name", sorted(ALL_CASES))
def test_oracle_matches_reference_bit_for_bit(name):
    g = np.load(os.path.join(GOLD, "ref_%s.npz" % name))
    s, _ = run_oracle(name)
    h = s.hist_array
    assert h.shape == g["history"].shape
    assert np.array_equal(h, g["history"])
    assert int(s.n_accepted) == int(g["n_accepted"])
    assert int(s.n_rejected) == int(g["n_rejected"])
    assert s.acceptance_fraction == float(g["acceptance_fraction"])
    if ALL_CASES[name]["algo"] == "dream":
        assert np.array_equal(s.p_cr, g["p_cr"])
        assert np.array_equal(s.delta_m, g["delta_m"])
        assert np.array_equal(s.n_cr_updates, g["n_cr_updates"])
    mean, std, _ = s.param_est(0)
    assert np.array_equal(mean, g["mean"]) and np.array_equal(std, g["std"])


@pytest.mark.parametrize("name", sorted(ALL_CASES))
def test_replay_form_matches_scalar_oracle(name):
    """Feeding the recorded draws to oracle/replay.py reproduces every generation's state
    and accept flags exactly, and p_cr at generation boundaries to rounding."""
    case = ALL_CASES[name]
    s, traces = run_oracle(name, record=True)
    fn, kw = oracle_target(case["target"])
    lnl = orp.scalar_batch(lambda th: fn(th, **kw))
    cfg = dict(algo=case["algo"], del_pairs=case["ctor_kwargs"].get("del_pairs", 3),
               n_cr=case["ctor_kwargs"].get("n_cr", 3),
               gamma_scale=case["ctor_kwargs"].get("gamma_scale", 1.0),
               gamma=case["run_kwargs"].get("gamma"),
               burnin_gen=case["ctor_kwargs"].get("burnin_gen", 300),
               n_cr_gen=case["ctor_kwargs"].get("n_cr_gen", 50))
    cr = orp.CrState(cfg["n_cr"]) if case["algo"] == "dream" else None
    hist = [s.history[0]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, tr in enumerate(traces):
            hv = None
            if case["algo"] == "dream" and len(hist) > cfg["n_cr_gen"]:
                hv = np.std(np.array(hist), axis=0) ** 2.0
            out = orp.replay_generation(hist[-1], tr, cfg, lnl, k, hist_var=hv, cr=cr)
            assert np.array_equal(out["state"], tr["state"]), "generation %d" % k
            assert np.array_equal(out["accept"], tr["accept"])
            assert np.array_equal(out["prop"], tr["prop"])
            if cr is not None:
                np.testing.assert_allclose(cr.p_cr, tr["p_cr"], rtol=1e-12)
                np.testing.assert_allclose(cr.delta_m, tr["delta_m"], rtol=1e-12)
                assert np.array_equal(cr.n_cr_updates, tr["n_cr_updates"])
            hist.append(out["state"])


def test_diagnostics_oracle_known_answers():
    """oracle/diagnostics.py against hand-computed values (the reference holds no vectors
    for the outlier reset / R-hat: parity for them is pinned to the published definitions)."""
    from oracle import diagnostics as od
    omega = np.array([-1.0, -2.0, -3.0, -4.0, -100.0])
    # Q1 = -4, Q3 = -2 -> threshold = -4 - 2*2 = -8
    mask, thr, best = od.iqr_outliers(omega)
    assert thr == -8.0 and best == 0 and mask.tolist() == [False, False, False, False, True]
    X = np.arange(10.0).reshape(5, 2)
    X2, L2, m = od.outlier_reset(X, omega.copy(), omega)
    assert X2[4].tolist() == [0.0, 1.0] and L2[4] == -1.0 and np.array_equal(X2[:4], X[:4])
    # two chains, identical within-chain variance 1, means 0 and 2, T = 3:
    h = np.array([[[-1.0], [1.0]], [[0.0], [2.0]], [[1.0], [3.0]]])
    # W = 1, B/T = var([0, 2], ddof=1) = 2 -> R = sqrt((2/3 + 2) / 1)
    np.testing.assert_allclose(od.rhat(h), [np.sqrt(2.0 / 3.0 + 2.0)], rtol=1e-15)


@pytest.mark.parametrize("name", ["serial_banana", "serial_gauss7"])
def test_serial_demc_oracle_matches_reference_bit_for_bit(name):
    """oracle/demc_serial.py vs the unmodified reference's DeMc (samplers.py:237-324)."""
    from oracle.cases import SERIAL_CASES, oracle_target
    from oracle.demc_serial import OracleDeMc
    case = SERIAL_CASES[name]
    g = np.load(os.path.join(GOLD, "ref_%s.npz" % name))
    fn, kw = oracle_target(case["target"])
    np.random.seed(case["seed"])
    s = OracleDeMc(fn, n_chains=case["n_chains"], ln_kwargs=kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        traces = s.run_mcmc(case["n"], case["theta_0"], record=True, **case["run_kwargs"])
    assert np.array_equal(np.array(s.history), g["history"])
    assert s.n_accepted == int(g["n_accepted"]) and s.n_rejected == int(g["n_rejected"])
    assert s.acceptance_fraction == float(g["acceptance_fraction"])
    mean, std, sl = s.param_est(0)
    assert np.array_equal(mean, g["mean"]) and np.array_equal(std, g["std"])
    assert np.array_equal(sl[:3 * case["n_chains"]], g["super_chain_head"])
    assert len(traces) == g["history"].shape[0] - 1
