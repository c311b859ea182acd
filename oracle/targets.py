"""CPU restatement of the reference's test likelihoods (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (bipymc_b200/) never does.

Each class restates the scalar ``ln_like`` of the matching reference class and keeps
its arithmetic route (scipy.stats.multivariate_normal ``pdf`` then ``np.log``), so
values are bit-identical to the reference's in this image (numpy 2.3 / scipy 1.18;
the pin is tests/test_oracle_golden.py against tests/golden/*.npz).

  Banana2D      <- bipymc/utils/banana_rv.py:10-37
  BimodeGauss2D <- bipymc/utils/dblgauss_rv.py:10-32
  GaussND       <- bipymc/utils/d100_gauss.py:10-35  (use_logpdf=True is the
                   1000-D capable variant SURVEY.md section 0 asks for)
  LineFit       <- examples/ex_para_fit.py:26-55 (lnprob of "EXAMPLE 2")
"""
import numpy as np
from scipy.stats import multivariate_normal


class Banana2D(object):
    def __init__(self, mu1=0, mu2=0, sigma1=1, sigma2=1, rho=0.9, a=1.15, b=0.5):
        self.a, self.b = a, b
        self.mean = np.array([mu1, mu2], dtype=float)
        self.cov = np.array([[sigma1 ** 2.0, rho * (sigma1 * sigma2)],
                             [rho * (sigma1 * sigma2), sigma2 ** 2.0]])
        self.rv = multivariate_normal(self.mean, self.cov)

    def pdf(self, y1, y2):
        x1 = y1 / self.a
        x2 = (y2 - self.b * (x1 ** 2.0 + self.a ** 2.0)) * self.a
        return self.rv.pdf(np.dstack((x1, x2)))

    def ln_like(self, y):
        assert len(y) == 2
        with np.errstate(divide="ignore"):
            return np.log(self.pdf(y[0], y[1]))

    def rvs(self, n):
        s = self.rv.rvs(size=n)
        x1, x2 = s[:, 0], s[:, 1]
        return self.a * x1, x2 / self.a + self.b * (x1 ** 2.0 + self.a ** 2.0)


class BimodeGauss2D(object):
    def __init__(self, mu_g1=(0, 0), mu_g2=(2, 2), sigma_g1=(0.25, 0.25), sigma_g2=(0.25, 0.25),
                 rho_g1=0.8, rho_g2=-0.8, w_g1=0.25, w_g2=0.75):
        self.mu_g1, self.mu_g2 = list(mu_g1), list(mu_g2)
        self.cov_g1 = np.array([[sigma_g1[0] ** 2.0, rho_g1 * (sigma_g1[0] * sigma_g1[1])],
                                [rho_g1 * (sigma_g1[0] * sigma_g1[1]), sigma_g1[1] ** 2.0]])
        self.cov_g2 = np.array([[sigma_g2[0] ** 2.0, rho_g2 * (sigma_g2[0] * sigma_g2[1])],
                                [rho_g2 * (sigma_g2[0] * sigma_g2[1]), sigma_g2[1] ** 2.0]])
        self.rv1 = multivariate_normal(self.mu_g1, self.cov_g1)
        self.rv2 = multivariate_normal(self.mu_g2, self.cov_g2)
        self.w_g1 = w_g1 / (w_g1 + w_g2)
        self.w_g2 = w_g2 / (w_g1 + w_g2)

    def pdf(self, y1, y2):
        pos = np.dstack((y1, y2))
        return self.w_g1 * self.rv1.pdf(pos) + self.w_g2 * self.rv2.pdf(pos)

    def ln_like(self, y):
        assert len(y) == 2
        with np.errstate(divide="ignore"):
            return np.log(self.pdf(y[0], y[1]))


def gauss_cov(dim, rho=0.5):
    """Sigma_ij = sqrt(i+1) sqrt(j+1) rho (i != j), Sigma_ii = i+1 (d100_gauss.py:17-25)."""
    sd = np.sqrt(np.arange(dim) + 1.0)
    cov = np.zeros((dim, dim))
    for i in range(dim):
        for j in range(dim):
            cov[i][j] = sd[i] ** 2.0 if i == j else sd[i] * sd[j] * rho
    return cov


class GaussND(object):
    def __init__(self, rho=0.5, dim=100, use_logpdf=False):
        self.dim = dim
        self.mu = np.zeros(dim)
        self.cov = gauss_cov(dim, rho)
        self.rv = multivariate_normal(self.mu, self.cov)
        self.use_logpdf = use_logpdf

    def ln_like(self, y):
        assert len(y) == self.dim
        if self.use_logpdf:
            return self.rv.logpdf(y)
        with np.errstate(divide="ignore"):
            return np.log(self.rv.pdf(y))


def linefit_data(seed=42, n=50, m_true=-0.9594, b_true=4.294, f_true=0.534):
    """Synthetic data of examples/ex_para_fit.py:17,26-35 (legacy RandomState stream)."""
    rs = np.random.RandomState(seed)
    x = np.sort(10 * rs.rand(n))
    yerr = 0.1 + 0.5 * rs.rand(n)
    y = m_true * x + b_true
    y += np.abs(f_true * y) * rs.randn(n)
    y += yerr * rs.randn(n)
    return x, y, yerr


class LineFit(object):
    """lnprob(theta=(m, b, lnf)) of examples/ex_para_fit.py:39-55."""
    def __init__(self, x=None, y=None, yerr=None):
        if x is None:
            x, y, yerr = linefit_data()
        self.x, self.y, self.yerr = np.asarray(x), np.asarray(y), np.asarray(yerr)

    def ln_like(self, theta):
        m, b, lnf = theta
        if not (-5.0 < m < 0.5 and 0.0 < b < 10.0 and -10.0 < lnf < 1.0):
            return -np.inf
        model = m * self.x + b
        inv_sigma2 = 1.0 / (self.yerr ** 2 + model ** 2 * np.exp(2 * lnf))
        return 0.0 + -0.5 * (np.sum((self.y - model) ** 2 * inv_sigma2 - np.log(inv_sigma2)))


def expfit_data(seed=42, n=60, tau=12.0, c_inf=1.5, c_0=0.6, leak=1e-3, sigma=2e-3):
    """Synthetic series for the model of examples/ex_exp_fit.py:38-43 (same generator as the
    product's bipymc_b200.targets.expfit_data; restated here so oracle/ stays self-contained)."""
    rs = np.random.RandomState(seed)
    t = np.linspace(0.5, 60.0, n)
    y = c_inf + c_0 * -1.0 * np.exp(-t / tau) - leak * t + np.sqrt(sigma) * 0.3 * rs.randn(n)
    return t, y


class ExpFit(object):
    """lnprob(theta, t, y_data) of examples/ex_exp_fit.py:73-121, function by function."""
    def __init__(self, t=None, y=None):
        if t is None:
            t, y = expfit_data()
        self.t, self.y = np.asarray(t), np.asarray(y)

    @staticmethod
    def exp_c1_model_full(tau, c_inf, c_0, leak, t):          # ex_exp_fit.py:38-43
        v2 = -1.0
        return c_inf + c_0 * v2 * np.exp(-t / tau) - leak * t

    @staticmethod
    def ln_params_prior(tau, c_inf, c_0, leak, sigma):        # ex_exp_fit.py:103-121
        if (-5 < c_inf < 5.0) and (1.0 < tau < 50) and (-1.0 < c_0 < 1.) and (-5. < leak < 5.0) \
                and (0 < sigma < 1.0):
            return 0.0
        return -np.inf

    def ln_like(self, theta):                                  # ex_exp_fit.py:73-101
        lp = self.ln_params_prior(*theta)
        if not np.isfinite(lp):
            return -np.inf
        tau, c_inf, c_0, leak, sigma = theta
        y_sigma = theta[-1]
        ln_model = np.sum((self.exp_c1_model_full(tau, c_inf, c_0, leak, self.t) - self.y) ** 2. / y_sigma
                          - np.log(1.0 / y_sigma))
        return lp + -0.5 * ln_model
