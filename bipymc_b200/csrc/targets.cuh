// Built-in batched log-likelihoods: the reference's test targets restated for the
// device.  Each follows the arithmetic route of the
// scalar reference (scipy.stats.multivariate_normal: eigh whitening, then
// -0.5 * (rank*log(2pi) + log_pdet + maha); the reference then takes log(exp(.)))
// so values agree with it to ~1e-13.
#pragma once
#include <stdint.h>
#include <math.h>

namespace bpm {

// 2-D frozen multivariate normal, scipy style: maha = |(x - mu) . U|^2.
struct Mvn2 {
  double mu0, mu1;
  double U00, U01, U10, U11;  // row-major prec_U (scipy _PSD.U)
  double c0;                  // rank * log(2 pi) + log_pdet
};
__device__ __forceinline__ double mvn2_logpdf(const Mvn2& g, double x0, double x1) {
  double d0 = __dsub_rn(x0, g.mu0), d1 = __dsub_rn(x1, g.mu1);
  double y0 = __dadd_rn(__dmul_rn(d0, g.U00), __dmul_rn(d1, g.U10));
  double y1 = __dadd_rn(__dmul_rn(d0, g.U01), __dmul_rn(d1, g.U11));
  double maha = __dadd_rn(__dmul_rn(y0, y0), __dmul_rn(y1, y1));
  return -0.5 * __dadd_rn(g.c0, maha);
}

// Host params layout (doubles): [log_of_pdf, a, b, mu0, mu1, U00, U01, U10, U11, c0]
struct BananaParams {
  double log_of_pdf, a, b;
  Mvn2 g;
};
// banana_rv.py:26-37
__device__ __forceinline__ double banana_lnl(const BananaParams& p, double y1, double y2) {
  double x1 = __ddiv_rn(y1, p.a);
  double t = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(p.a, p.a));
  double x2 = __dmul_rn(__dsub_rn(y2, __dmul_rn(p.b, t)), p.a);
  double lp = mvn2_logpdf(p.g, x1, x2);
  return p.log_of_pdf != 0.0 ? log(exp(lp)) : lp;
}

// Host params layout: [log_of_pdf, w1, w2, mvn2 #1 (7), mvn2 #2 (7)]
struct BimodalParams {
  double log_of_pdf, w1, w2;
  Mvn2 g1, g2;
};
// dblgauss_rv.py:26-32
__device__ __forceinline__ double bimodal_lnl(const BimodalParams& p, double y1, double y2) {
  double l1 = mvn2_logpdf(p.g1, y1, y2), l2 = mvn2_logpdf(p.g2, y1, y2);
  if (p.log_of_pdf != 0.0)
    return log(__dadd_rn(__dmul_rn(p.w1, exp(l1)), __dmul_rn(p.w2, exp(l2))));
  double a = l1 + log(p.w1), b = l2 + log(p.w2);
  double m = fmax(a, b);
  if (!(m > -INFINITY)) return -INFINITY;
  return m + log(exp(a - m) + exp(b - m));
}

// Gaussian: lnl = -0.5 * (c0 + |(x - mu) . W|^2), optional log(exp(.)) round trip
// (d100_gauss.py:29-35 takes np.log(pdf)).
__device__ __forceinline__ double gauss_finish(double c0, double maha, int log_of_pdf) {
  double lp = -0.5 * __dadd_rn(c0, maha);
  return log_of_pdf ? log(exp(lp)) : lp;
}

// Line fit, examples/ex_para_fit.py:39-55.  data = x[M], y[M], yerr[M] (device).
// The M terms are independent up to the (ordered) accumulation: unrolled so that five divisions / logarithms
// are in flight per thread.  (Measured alternatives, profiles/r2/r2r_secondary.txt, r2s_*: four interleaved
// partial sums 76 us per generation of 10^5 chains, this form 73 us, four LANES per chain 86 us.)
__device__ __forceinline__ double linefit_lnl(const double* x, const double* y, const double* yerr,
                                              int M, double m, double b, double lnf) {
  if (!(-5.0 < m && m < 0.5 && 0.0 < b && b < 10.0 && -10.0 < lnf && lnf < 1.0)) return -INFINITY;
  const double e2 = exp(__dmul_rn(2.0, lnf));
  double s = 0.0;
#pragma unroll 5
  for (int i = 0; i < M; ++i) {
    double model = __dadd_rn(__dmul_rn(m, x[i]), b);
    double inv = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(yerr[i], yerr[i]),
                                          __dmul_rn(__dmul_rn(model, model), e2)));
    double r = __dsub_rn(y[i], model);
    s += __dsub_rn(__dmul_rn(__dmul_rn(r, r), inv), log(inv));
  }
  return 0.0 + -0.5 * s;
}

// Exponential-relaxation fit, examples/ex_exp_fit.py:38-121.
//   theta = (tau, c_inf, c_0, leak, sigma);  flat box prior (ex_exp_fit.py:103-121);
//   model(t) = c_inf + c_0 * (-1) * exp(-t / tau) - leak * t          (ex_exp_fit.py:38-43)
//   lnprob   = 0.0 + -0.5 * sum_i [ (model_i - y_i)^2 / sigma - log(1 / sigma) ]   (ex_exp_fit.py:73-101)
__device__ __forceinline__ double expfit_lnl(const double* t, const double* y, int M, double tau, double c_inf,
                                             double c_0, double leak, double sigma) {
  if (!(-5.0 < c_inf && c_inf < 5.0 && 1.0 < tau && tau < 50.0 && -1.0 < c_0 && c_0 < 1.0 && -5.0 < leak &&
        leak < 5.0 && 0.0 < sigma && sigma < 1.0))
    return -INFINITY;
  const double c0v = __dmul_rn(c_0, -1.0);
  const double lg = log(__ddiv_rn(1.0, sigma));
  double s = 0.0;
  for (int i = 0; i < M; ++i) {
    const double ex = exp(__ddiv_rn(-t[i], tau));
    const double m = __dsub_rn(__dadd_rn(c_inf, __dmul_rn(c0v, ex)), __dmul_rn(leak, t[i]));
    const double r = __dsub_rn(m, y[i]);
    s += __dsub_rn(__ddiv_rn(__dmul_rn(r, r), sigma), lg);
  }
  return 0.0 + -0.5 * s;
}

}  // namespace bpm
