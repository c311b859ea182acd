"""IQR outlier-chain reset and Gelman-Rubin R-hat through the C-ABI (bpm_outlier_reset,
bpm_rhat, bpm_omega_track) against oracle/diagnostics.py: WHICH chains are reset, the
threshold and the copied states must agree exactly; R-hat to 1e-10."""
import ctypes as C

import numpy as np
import pytest

from oracle import diagnostics as odiag

pytestmark = pytest.mark.gpu


def _sampler(n=64, **kw):
    from bipymc_b200 import DreamMpi, targets
    np.random.seed(5)
    return DreamMpi(targets.BimodeGauss_2D().ln_like, [0.0, 0.0], n_chains=n, seed=11, varepsilon=0.3,
                    burnin_gen=10000, n_cr_gen=5, **kw)


@pytest.mark.parametrize("n", [8, 64, 1001, 100000])
def test_outlier_reset_matches_oracle_on_explicit_omega(n):
    import torch
    from bipymc_b200 import _lib
    s = _sampler(n)
    s.run_mcmc(n * 3)
    rs = np.random.RandomState(n)
    omega = rs.standard_normal(n) * 3.0 - 10.0
    omega[rs.choice(n, size=max(1, n // 16), replace=False)] -= 40.0     # planted outliers
    if n == 64:
        omega[3] = -np.inf                                               # a chain stuck at -inf
    X = s._X[:, :2].cpu().numpy()
    lnl = s._lnl.cpu().numpy()
    wantX, wantL, mask = odiag.outlier_reset(X, lnl, omega)
    om = torch.from_numpy(omega).to(s._device)
    flags = torch.zeros(n, dtype=torch.int32, device=s._device)
    n_reset, stats = C.c_int32(), (C.c_double * 4)()
    st = s._state(None)
    _lib.check(s._libh.bpm_outlier_reset(s._handle, C.byref(st), om.data_ptr(), flags.data_ptr(),
                                         C.byref(n_reset), stats, s._stream()))
    _, thr, best = odiag.iqr_outliers(omega)
    assert stats[0] == thr and int(stats[3]) == best
    assert np.array_equal(flags.cpu().numpy().astype(bool), mask)
    assert n_reset.value == int(mask.sum()) and n_reset.value >= 1
    assert np.array_equal(s._X[:, :2].cpu().numpy(), wantX)
    assert np.array_equal(s._lnl.cpu().numpy(), wantL)


def test_no_outliers_means_no_change():
    import torch
    from bipymc_b200 import _lib
    s = _sampler(32)
    s.run_mcmc(32 * 2)
    before = s._X.clone()
    om = torch.full((32,), -3.0, dtype=torch.float64, device=s._device)   # IQR = 0, nobody below
    n_reset = C.c_int32(7)
    st = s._state(None)
    _lib.check(s._libh.bpm_outlier_reset(s._handle, C.byref(st), om.data_ptr(), None, C.byref(n_reset),
                                         None, s._stream()))
    assert n_reset.value == 0 and torch.equal(before, s._X)


def test_tracked_omega_is_the_mean_cached_loglike_and_reset_runs_in_burnin():
    """bpm_omega_track sums the cached ln_like once per generation; run_mcmc(outlier_gen=K)
    checks every K generations while k < burnin_gen.  A chain parked in the tail of the
    bimodal target is an outlier at the first check and must be pulled back."""
    import torch
    from bipymc_b200 import _lib
    n, K = 48, 10
    s = _sampler(n, outlier_gen=K)
    s._X[5, :2] = torch.tensor([3.4, 3.4], dtype=torch.float64, device=s._device)
    s._lnl_valid = False
    s.run_mcmc(n * (K + 1))                       # exactly K generations -> one check
    assert s.n_outlier_resets >= 1
    assert s.last_outlier_stats["threshold"] < s.last_outlier_stats["q1"]
    assert float(s._X[5, :2].abs().max()) < 3.2
    # Omega sums restarted after the check
    cnt, p = C.c_int64(), C.c_void_p()
    _lib.check(s._libh.bpm_omega(s._handle, C.byref(p), C.byref(cnt)))
    assert cnt.value == 0
    # explicit window: sums == sum of the lnl history rows
    s2 = _sampler(n, outlier_gen=1000)
    rows = []
    for g in range(4):
        s2.run_mcmc(2 * n, _k_gen0=g) if g else s2.run_mcmc(2 * n)
        rows.append(s2._lnl.cpu().numpy().copy())
        _lib.check(s2._libh.bpm_omega(s2._handle, C.byref(p), C.byref(cnt)))
        assert cnt.value == 1                      # every run_mcmc call restarts the window
        got = s2._wrap_device(p.value, (n,)).cpu().numpy()
        assert np.array_equal(got, rows[-1])


def test_rhat_history_and_streaming_match_oracle():
    from bipymc_b200 import DreamMpi, targets
    np.random.seed(2)
    s = DreamMpi(targets.Gauss_100D(dim=12).ln_like, np.zeros(12), n_chains=200, seed=9, varepsilon=1.0)
    s.run_mcmc(200 * 301)
    h = s._hist.tensor()[:, :, :12].cpu().numpy()
    np.testing.assert_allclose(s.rhat(), odiag.rhat(h), rtol=1e-10)
    np.testing.assert_allclose(s.rhat_history(), odiag.rhat(h[h.shape[0] // 2:]), rtol=1e-10)
    np.testing.assert_allclose(s.rhat_history(100), odiag.rhat(h[100:]), rtol=1e-10)
    with pytest.raises(Exception):
        s.rhat_history(10 ** 6)
