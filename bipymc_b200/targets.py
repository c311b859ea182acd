"""Built-in likelihood targets with a batched device implementation.

Mirrors the reference's test-target classes (same constructor arguments, same ``pdf`` /
``ln_like`` / ``rvs`` surface):

  Banana_2D       bipymc/utils/banana_rv.py:10-61
  BimodeGauss_2D  bipymc/utils/dblgauss_rv.py:10-43
  Gauss_100D      bipymc/utils/d100_gauss.py:10-39
  LineFit         examples/ex_para_fit.py:26-55 (lnprob of the second example)
  ExpFit          examples/ex_exp_fit.py:38-121 (5-parameter relaxation fit)

``ln_like`` stays a scalar host function, so the objects remain valid ``ln_like_fn``
arguments everywhere; when a sampler of this package receives ``obj.ln_like`` it
recognises the owner through ``device_target()`` and evaluates the likelihood inside the
CUDA generation kernels instead (bipymc_b200/csrc/targets.cuh).

Host-side parameter preparation follows scipy.stats.multivariate_normal (which the
reference calls): symmetric eigendecomposition, ``U = V / sqrt(lambda)``,
``maha = |(x - mu) U|^2``, ``logpdf = -0.5 (rank log 2pi + log pdet + maha)``
(scipy/stats/_multivariate.py ``_PSD`` and ``_logpdf``).
"""
import numpy as np

from . import _lib

_LOG_2PI = np.log(2 * np.pi)


def psd_whiten(cov):
    """(U, log_pdet, rank) of a covariance, the way scipy's _PSD computes them."""
    cov = np.asarray(cov, dtype=float)
    s, u = np.linalg.eigh(cov)
    eps = 1e6 * np.finfo("d").eps * np.max(np.abs(s))     # scipy's _eigvalsh_to_eps default
    if np.min(s) < -eps:
        raise ValueError("the input matrix must be symmetric positive semidefinite")
    keep = s > eps
    d = s[keep]
    s_pinv = np.array([0.0 if abs(x) <= eps else 1.0 / x for x in s])
    U = np.multiply(u, np.sqrt(s_pinv))
    U = U[:, keep]
    return np.ascontiguousarray(U), float(np.sum(np.log(d))), int(len(d))


class _Mvn(object):
    """Minimal frozen multivariate normal (pdf / logpdf / rvs) on the host."""
    def __init__(self, mean, cov):
        self.mean = np.asarray(mean, dtype=float)
        self.cov = np.asarray(cov, dtype=float)
        self.U, self.log_pdet, self.rank = psd_whiten(self.cov)
        self.c0 = self.rank * _LOG_2PI + self.log_pdet

    def logpdf(self, x):
        dev = np.asarray(x, dtype=float) - self.mean
        maha = np.sum(np.square(np.dot(dev, self.U)), axis=-1)
        return -0.5 * (self.c0 + maha)

    def pdf(self, x):
        return np.exp(self.logpdf(x))

    def rvs(self, size=1):
        return np.random.multivariate_normal(self.mean, self.cov, size=size)

    def packed2(self):
        assert self.mean.shape == (2,) and self.U.shape == (2, 2)
        return [self.mean[0], self.mean[1], self.U[0, 0], self.U[0, 1], self.U[1, 0], self.U[1, 1],
                self.c0]


class DeviceTarget(object):
    """(target id, flat float64 parameter array) handed to bpm_set_target."""
    def __init__(self, target_id, params, dim):
        self.target_id = int(target_id)
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        self.dim = int(dim)


class Banana_2D(object):
    def __init__(self, mu1=0, mu2=0, sigma1=1, sigma2=1, rho=0.9, a=1.15, b=0.5, log_of_pdf=True):
        self.mu1, self.mu2 = mu1, mu2
        self.sigma1, self.sigma2, self.rho = sigma1, sigma2, rho
        self.a, self.b = a, b
        self.log_of_pdf = log_of_pdf
        cov = np.array([[sigma1 ** 2.0, rho * (sigma1 * sigma2)],
                        [rho * (sigma1 * sigma2), sigma2 ** 2.0]])
        self.rv_2d_normal = _Mvn(np.array([mu1, mu2], dtype=float), cov)

    def inv_transform(self, y1, y2):
        x1 = y1 / self.a
        x2 = (y2 - self.b * (x1 ** 2.0 + self.a ** 2.0)) * self.a
        return x1, x2

    def transform(self, x1, x2):
        return self.a * x1, x2 / self.a + self.b * (x1 ** 2.0 + self.a ** 2.0)

    def pdf(self, y1, y2):
        x1, x2 = self.inv_transform(np.asarray(y1, dtype=float), np.asarray(y2, dtype=float))
        return self.rv_2d_normal.pdf(np.stack((x1, x2), axis=-1))

    def ln_like(self, y):
        assert len(y) == 2
        with np.errstate(divide="ignore"):
            return np.log(self.pdf(y[0], y[1]))

    def check_prob_lvl(self, y1, y2, pdf_lvl):
        return pdf_lvl < self.pdf(y1, y2)

    def rvs(self, n_samples):
        s = self.rv_2d_normal.rvs(size=n_samples)
        return self.transform(s[:, 0], s[:, 1])

    def device_target(self):
        p = [1.0 if self.log_of_pdf else 0.0, self.a, self.b] + self.rv_2d_normal.packed2()
        return DeviceTarget(_lib.TARGET_BANANA, p, 2)


class BimodeGauss_2D(object):
    def __init__(self, mu_g1=[0, 0], mu_g2=[2, 2], sigma_g1=[0.25, 0.25], sigma_g2=[0.25, 0.25],
                 rho_g1=0.8, rho_g2=-0.8, w_g1=0.25, w_g2=0.75, log_of_pdf=True):
        self.mu_g1, self.mu_g2 = mu_g1, mu_g2
        self.cov_g1 = np.array([[sigma_g1[0] ** 2.0, rho_g1 * (sigma_g1[0] * sigma_g1[1])],
                                [rho_g1 * (sigma_g1[0] * sigma_g1[1]), sigma_g1[1] ** 2.0]])
        self.cov_g2 = np.array([[sigma_g2[0] ** 2.0, rho_g2 * (sigma_g2[0] * sigma_g2[1])],
                                [rho_g2 * (sigma_g2[0] * sigma_g2[1]), sigma_g2[1] ** 2.0]])
        self.rv_2d_g1 = _Mvn(self.mu_g1, self.cov_g1)
        self.rv_2d_g2 = _Mvn(self.mu_g2, self.cov_g2)
        self.w_g1 = w_g1 / (w_g1 + w_g2)
        self.w_g2 = w_g2 / (w_g1 + w_g2)
        self.log_of_pdf = log_of_pdf

    def pdf(self, y1, y2):
        pos = np.stack((np.asarray(y1, dtype=float), np.asarray(y2, dtype=float)), axis=-1)
        return self.w_g1 * self.rv_2d_g1.pdf(pos) + self.w_g2 * self.rv_2d_g2.pdf(pos)

    def ln_like(self, y):
        assert len(y) == 2
        if not self.log_of_pdf:
            # direct log-density (log-sum-exp): finite far from both modes, where the
            # reference's np.log(pdf) underflows to -inf
            pos = np.asarray(y, dtype=float)
            a = self.rv_2d_g1.logpdf(pos) + np.log(self.w_g1)
            b = self.rv_2d_g2.logpdf(pos) + np.log(self.w_g2)
            m = max(a, b)
            return m + np.log(np.exp(a - m) + np.exp(b - m))
        with np.errstate(divide="ignore"):
            return np.log(self.pdf(y[0], y[1]))

    def rvs(self, n_samples):
        sel = np.random.choice((True, False), p=(self.w_g1, self.w_g2), size=n_samples)
        out = np.zeros((n_samples, 2))
        s1, s2 = self.rv_2d_g1.rvs(size=n_samples), self.rv_2d_g2.rvs(size=n_samples)
        out[sel, :] = s1[sel, :]
        out[~sel, :] = s2[~sel, :]
        return out[:, 0], out[:, 1]

    def device_target(self):
        p = ([1.0 if self.log_of_pdf else 0.0, self.w_g1, self.w_g2] + self.rv_2d_g1.packed2() +
             self.rv_2d_g2.packed2())
        return DeviceTarget(_lib.TARGET_BIMODAL, p, 2)


def gauss_cov(dim, rho=0.5):
    """d100_gauss.py:17-25: var_i = i + 1, pairwise correlation rho."""
    sd = np.sqrt(np.arange(dim) + 1.0)
    cov = np.outer(sd, sd) * rho
    cov[np.diag_indices(dim)] = sd ** 2.0
    return cov


class Gauss_100D(object):
    """Correlated Gaussian of any dimension.  ``log_of_pdf=True`` reproduces the
    reference's ``np.log(pdf)`` (which underflows to -inf beyond d of a few hundred);
    ``log_of_pdf=False`` evaluates the log-density directly and is what dim=1000 needs.
    A general (mean, cov) can be passed instead of the rho rule."""
    def __init__(self, rho=0.5, dim=100, log_of_pdf=None, mean=None, cov=None):
        self.dim = dim
        self.rho = rho
        self.mu = np.zeros(dim) if mean is None else np.asarray(mean, dtype=float)
        self.var = np.sqrt(np.arange(dim) + 1.0)
        self.cov = gauss_cov(dim, rho) if cov is None else np.asarray(cov, dtype=float)
        self.rv_100d = _Mvn(self.mu, self.cov)
        self.log_of_pdf = (dim <= 200) if log_of_pdf is None else bool(log_of_pdf)

    def pdf(self, y):
        return self.rv_100d.pdf(y)

    def ln_like(self, y):
        assert len(y) == self.dim
        if not self.log_of_pdf:
            return self.rv_100d.logpdf(y)
        with np.errstate(divide="ignore"):
            return np.log(self.pdf(y))

    def rvs(self, n_samples):
        return self.rv_100d.rvs(size=n_samples)

    def device_target(self):
        rv = self.rv_100d
        head = [1.0 if self.log_of_pdf else 0.0, rv.c0, float(rv.U.shape[1])]
        return DeviceTarget(_lib.TARGET_GAUSS, np.concatenate([head, rv.mean, rv.U.ravel()]), self.dim)


def linefit_data(seed=42, n=50, m_true=-0.9594, b_true=4.294, f_true=0.534):
    """Synthetic data of examples/ex_para_fit.py:17,26-35."""
    rs = np.random.RandomState(seed)
    x = np.sort(10 * rs.rand(n))
    yerr = 0.1 + 0.5 * rs.rand(n)
    y = m_true * x + b_true
    y += np.abs(f_true * y) * rs.randn(n)
    y += yerr * rs.randn(n)
    return x, y, yerr


class LineFit(object):
    """theta = (m, b, ln f): box prior + heteroscedastic Gaussian likelihood."""
    def __init__(self, x=None, y=None, yerr=None):
        if x is None:
            x, y, yerr = linefit_data()
        self.x, self.y, self.yerr = (np.asarray(v, dtype=float) for v in (x, y, yerr))

    def ln_like(self, theta):
        m, b, lnf = theta
        if not (-5.0 < m < 0.5 and 0.0 < b < 10.0 and -10.0 < lnf < 1.0):
            return -np.inf
        model = m * self.x + b
        inv_sigma2 = 1.0 / (self.yerr ** 2 + model ** 2 * np.exp(2 * lnf))
        return 0.0 + -0.5 * (np.sum((self.y - model) ** 2 * inv_sigma2 - np.log(inv_sigma2)))

    def device_target(self):
        p = np.concatenate([[float(len(self.x))], self.x, self.y, self.yerr])
        return DeviceTarget(_lib.TARGET_LINEFIT, p, 3)


def expfit_data(seed=42, n=60, tau=12.0, c_inf=1.5, c_0=0.6, leak=1e-3, sigma=2e-3):
    """Synthetic relaxation data for the model of examples/ex_exp_fit.py:38-43 (the example
    reads a measured .mat series; a seeded synthetic series of the same shape stands in)."""
    rs = np.random.RandomState(seed)
    t = np.linspace(0.5, 60.0, n)
    y = c_inf + c_0 * -1.0 * np.exp(-t / tau) - leak * t + np.sqrt(sigma) * 0.3 * rs.randn(n)
    return t, y


class ExpFit(object):
    """theta = (tau, c_inf, c_0, leak, sigma): lnprob of examples/ex_exp_fit.py:73-121."""
    def __init__(self, t=None, y=None):
        if t is None:
            t, y = expfit_data()
        self.t, self.y = np.asarray(t, dtype=float), np.asarray(y, dtype=float)

    def ln_like(self, theta):
        tau, c_inf, c_0, leak, sigma = theta
        if not (-5.0 < c_inf < 5.0 and 1.0 < tau < 50.0 and -1.0 < c_0 < 1.0 and -5.0 < leak < 5.0 and
                0.0 < sigma < 1.0):
            return -np.inf
        m = c_inf + c_0 * -1.0 * np.exp(-self.t / tau) - leak * self.t
        return 0.0 + -0.5 * np.sum((m - self.y) ** 2. / sigma - np.log(1.0 / sigma))

    def device_target(self):
        p = np.concatenate([[float(len(self.t))], self.t, self.y])
        return DeviceTarget(_lib.TARGET_EXPFIT, p, 5)


def resolve_device_target(ln_like_fn, ln_kwargs):
    """Return a DeviceTarget when ``ln_like_fn`` is the ``ln_like`` of a built-in target
    (or the object itself) and no extra kwargs are bound; otherwise None."""
    if ln_kwargs:
        return None
    owner = getattr(ln_like_fn, "__self__", None)
    if owner is not None and getattr(ln_like_fn, "__name__", "") == "ln_like" and \
            hasattr(owner, "device_target"):
        return owner.device_target()
    if hasattr(ln_like_fn, "device_target") and not callable(getattr(ln_like_fn, "__call__", None)):
        return ln_like_fn.device_target()
    if isinstance(ln_like_fn, DeviceTarget):
        return ln_like_fn
    return None
