"""Serial DeMc (bipymc/samplers.py:237-324, delayed accept) on the device.
(1) RNG replay: the reference's recorded numpy draws reproduce its chains (golden vectors from
    the unmodified reference) step for step: accept flags equal, states <= 1e-12 relative.
(2) native Philox draws, dumped and replayed by the oracle, land on the same states.
(3) the reference's posterior gate for the banana (tests/test_banana.py:66-72)."""
import os
import warnings

import numpy as np
import pytest

from oracle.cases import SERIAL_CASES, oracle_target
from oracle.demc_serial import OracleDeMc, replay_serial_generation
from oracle import targets as otargets

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-12


def _device_target(name):
    from bipymc_b200 import targets
    return targets.Banana_2D(sigma1=1.0, sigma2=1.0) if name == "banana" else targets.Gauss_100D(dim=int(name[5:]))


@pytest.mark.parametrize("mode", ["device", "scalar"])
@pytest.mark.parametrize("name", sorted(SERIAL_CASES))
def test_serial_replay_matches_reference(name, mode):
    from bipymc_b200 import DeMc
    case = SERIAL_CASES[name]
    g = np.load(os.path.join(GOLD, "ref_%s.npz" % name))
    fn, kw = oracle_target(case["target"])
    np.random.seed(case["seed"])
    o = OracleDeMc(fn, n_chains=case["n_chains"], ln_kwargs=kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        traces = o.run_mcmc(case["n"], case["theta_0"], record=True, **case["run_kwargs"])
    tgt = _device_target(case["target"])
    like = tgt.ln_like if mode == "device" else (lambda th: tgt.ln_like(th))
    np.random.seed(case["seed"])
    s = DeMc(like, n_chains=case["n_chains"], seed=3)
    sink = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s.run_mcmc(case["n"], np.asarray(case["theta_0"], dtype=float), _replay=traces,
                   _trace=sink if mode == "device" else None, **case["run_kwargs"])
    assert s._mode() == mode
    hist = s._hist.tensor()[:, :, :s.dim].cpu().numpy()
    ref = g["history"]
    assert hist.shape == ref.shape and np.array_equal(hist[0], ref[0])
    err = np.abs(hist - ref) / np.maximum(1.0, np.abs(ref))
    assert err.max() <= RTOL, err.max()
    for got, tr in zip(sink, traces):
        assert np.array_equal(got["accept"], tr["accept"])
    assert s.n_accepted == int(g["n_accepted"]) and s.n_rejected == int(g["n_rejected"])
    assert s.acceptance_fraction == float(g["acceptance_fraction"])
    mean, std, sl = s.param_est(0)
    np.testing.assert_allclose(mean, g["mean"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(sl[:3 * case["n_chains"]], g["super_chain_head"], rtol=1e-12, atol=1e-13)
    with pytest.raises(NotImplementedError):
        s.run_mcmc(100, np.asarray(case["theta_0"], dtype=float), delayed_accept=False)


def test_serial_native_matches_oracle_replay():
    from bipymc_b200 import DeMc, targets
    N, d = 40, 12
    tgt, otgt = targets.Gauss_100D(dim=d), otargets.GaussND(dim=d)
    np.random.seed(2)
    s = DeMc(tgt.ln_like, n_chains=N, seed=77)
    s.run_mcmc(N * 3, np.zeros(d), varepsilon=1e-2)               # initialise + two sweeps
    gamma = 2.38 / np.sqrt(2.0 * d)
    for k in range(4):
        tr = s._dump_native_draws(k)
        pre = s._X[:, :d].cpu().numpy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want, acc, _ = replay_serial_generation(pre, tr, gamma, otgt.ln_like)
        st = s._state(s._hist.reserve(1)[0])
        import ctypes as C
        from bipymc_b200 import _lib
        _lib.check(s._libh.bpm_reset_counters(s._handle))
        _lib.check(s._libh.bpm_step_generations(s._handle, C.byref(st), k, 1, s._stream()))
        s._hist.advance(1)
        s._mom_len += 1
        got = s._X[:, :d].cpu().numpy()
        err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
        assert err.max() <= RTOL, (k, err.max())
        a, r = C.c_uint64(), C.c_uint64()
        _lib.check(s._libh.bpm_get_counters(s._handle, C.byref(a), C.byref(r), None))
        assert a.value == int(acc.sum()) and a.value + r.value == N


def test_serial_banana_posterior_gate():
    """tests/test_banana.py:66-72 applied to the serial sampler: fraction of samples inside
    two probability levels within 0.05 of the truth."""
    from bipymc_b200 import DeMc, targets
    banana = targets.Banana_2D()
    np.random.seed(42)
    y1, y2 = banana.rvs(200000)
    p = banana.pdf(y1, y2)
    want = [(p > 0.18).mean(), (p > 0.018).mean()]
    s = DeMc(banana.ln_like, n_chains=64, seed=5)
    s.run_mcmc(64 * 2001, np.array([0.0, 0.0]), varepsilon=1e-3)
    _, _, sl = s.param_est(64 * 500)
    ps = banana.pdf(sl[:, 0], sl[:, 1])
    got = [(ps > 0.18).mean(), (ps > 0.018).mean()]
    assert abs(got[0] - want[0]) < 0.05 and abs(got[1] - want[1]) < 0.05, (got, want)
    assert 0.05 < s.acceptance_fraction < 0.8
