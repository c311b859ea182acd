// Burn-in diagnostics the north star adds to the reference's path (the reference has no
// counterpart; definitions from Vrugt et al. 2009, the paper bipymc cites at readme.md:41-43):
//   * IQR outlier-chain reset: Omega_c = mean log-density of chain c over a trailing window;
//     chains with Omega_c < Q1 - 2 (Q3 - Q1) jump to the current state of the best chain.
//     Quartiles follow numpy.percentile's default (linear) rule so that the CPU oracle
//     (oracle/diagnostics.py) and the device agree on WHICH chains are reset, bit for bit.
//   * Gelman-Rubin R-hat per dimension from per-chain means / sums of squared deviations
//     (either the running Welford moments or moments of stored history rows [t0, T)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bpm {

// Omega bookkeeping: sum of the cached log-likelihood of every local chain, one add per
// generation (N doubles of traffic against ~10 kB per chain-step of the update itself).
__global__ void omega_accum_kernel(const double* __restrict__ lnl, double* __restrict__ omega_sum, int lo,
                                   int hi) {
  const int c = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (c < hi) omega_sum[c] += lnl[c];
}

__global__ void omega_mean_kernel(const double* __restrict__ omega_sum, double inv_cnt, int n,
                                  double* __restrict__ omega) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) omega[c] = omega_sum[c] * inv_cnt;
}

// numpy.percentile(x, q) with the default linear method on an ascending array
// (numpy/lib/_function_base_impl.py _lerp): a + (b - a) t, or b - (b - a)(1 - t) for t >= 0.5.
__device__ __forceinline__ double np_quantile_sorted(const double* __restrict__ s, int n, double q) {
  const double pos = __dmul_rn(q, (double)(n - 1));   // exact for q in {0.25, 0.75}
  int lo = (int)floor(pos);
  lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
  const int hi = lo + 1 > n - 1 ? n - 1 : lo + 1;
  const double t = __dsub_rn(pos, (double)lo);
  const double a = s[lo], b = s[hi];
  const double diff = __dsub_rn(b, a);
  if (t >= 0.5) return __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  return __dadd_rn(a, __dmul_rn(diff, t));
}

// out[0] = threshold Q1 - 2 IQR, out[1] = Q1, out[2] = Q3
__global__ void iqr_threshold_kernel(const double* __restrict__ sorted, int n, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double q1 = np_quantile_sorted(sorted, n, 0.25);
  const double q3 = np_quantile_sorted(sorted, n, 0.75);
  const double iqr = __dsub_rn(q3, q1);
  out[0] = __dsub_rn(q1, __dmul_rn(2.0, iqr));
  out[1] = q1;
  out[2] = q3;
}

// best chain = argmax Omega, ties to the lowest id (numpy argmax); one block.
__global__ void __launch_bounds__(1024) argmax_kernel(const double* __restrict__ v, int n,
                                                      int32_t* __restrict__ best) {
  __shared__ double sv[32];
  __shared__ int si[32];
  double bv = -INFINITY;
  int bi = 0x7FFFFFFF;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = v[i];
    if (x > bv || (x == bv && i < bi)) { bv = x; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
    const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7FFFFFFF;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
      const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) *best = bi == 0x7FFFFFFF ? 0 : bi;   // all -inf / NaN: chain 0
  }
}

// One warp per local chain: outliers copy the best chain's row, cached lnL and Omega sum.
__global__ void __launch_bounds__(256) outlier_reset_kernel(double* __restrict__ X, double* __restrict__ lnl,
                                                            double* __restrict__ omega_sum,
                                                            const double* __restrict__ omega,
                                                            const double* __restrict__ thr,
                                                            const int32_t* __restrict__ best, int lo, int hi,
                                                            int d, int ld, int32_t* __restrict__ flags,
                                                            int32_t* __restrict__ n_reset) {
  const int c = lo + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (c >= hi) return;
  const int b = *best;
  const bool out = omega[c] < thr[0] && c != b;
  if (flags && lane == 0) flags[c] = out ? 1 : 0;
  if (!out) return;
  const double* src = X + (size_t)b * ld;
  double* dst = X + (size_t)c * ld;
  for (int i = lane; i < d; i += 32) dst[i] = src[i];
  if (lane == 0) {
    lnl[c] = lnl[b];
    if (omega_sum) omega_sum[c] = omega_sum[b];
    atomicAdd(n_reset, 1);
  }
}

// Per-(chain, dimension) mean and sum of squared deviations of history rows [t0, T).
__global__ void history_moments_kernel(const double* __restrict__ hist, int64_t t0, int64_t T, int n_local,
                                       int d, int ld, double* __restrict__ mean, double* __restrict__ m2) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n_local * ld) return;
  const int i = (int)(idx % ld);
  if (i >= d) return;
  double mu = 0.0, s2 = 0.0;
  for (int64_t t = t0; t < T; ++t) {
    const double s = hist[(size_t)t * n_local * ld + idx];
    const double dl = s - mu;
    mu += dl / (double)(t - t0 + 1);
    s2 += dl * (s - mu);
  }
  mean[idx] = mu;
  m2[idx] = s2;
}

// R-hat of dimension blockIdx.x from n chains' (mean, m2) over `rows` rows each:
//   W = mean_c m2_c / (rows - 1),  B/rows = var_c(mean_c) (ddof 1),
//   R = sqrt(((rows - 1) / rows W + B/rows) / W)           (Gelman & Rubin 1992)
// Two fixed-order passes over the chains (mean of means first), so the result does not
// depend on the launch geometry.
__global__ void __launch_bounds__(256) rhat_kernel(const double* __restrict__ mean, const double* __restrict__ m2,
                                                   int n, int ld, double rows, double* __restrict__ rhat) {
  __shared__ double sa[256], sb[256];
  __shared__ double gm;
  const int i = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    a += mean[(size_t)c * ld + i];
    b += m2[(size_t)c * ld + i];
  }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) gm = sa[0] / (double)n;
  const double W = sb[0] / (double)n / (rows - 1.0);
  __syncthreads();
  double v = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const double dm = mean[(size_t)c * ld + i] - gm;
    v += dm * dm;
  }
  __syncthreads();
  sa[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sa[threadIdx.x] += sa[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double B_over_T = sa[0] / (double)(n - 1);
    rhat[i] = sqrt(((rows - 1.0) / rows * W + B_over_T) / W);
  }
}

// Streaming cross moments of the population (posterior COVARIANCE without a stored history: north_star's
// "posterior mean and covariance ... within Monte-Carlo standard error"): after every generation
//   sum[i] += sum_c X[c][i],   cross[i][j] += sum_c X[c][i] X[c][j]   over this rank's chains.
// O(n_local d^2) per generation -- a diagnostic for populations of 10^2 .. 10^4 chains, off by default.
// One block per row i of the d x d matrix; thread j walks the chains (X[c][i] is a broadcast, X[c][j] coalesced).
__global__ void __launch_bounds__(128) cross_moment_kernel(const double* __restrict__ X, int lo, int hi, int d, int ld,
                                                           double* __restrict__ sum, double* __restrict__ cross) {
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    double acc = 0.0, s1 = 0.0;
    for (int c = lo; c < hi; ++c) {
      const double xi = X[(size_t)c * ld + i], xj = X[(size_t)c * ld + j];
      acc = fma(xi, xj, acc);
      s1 += xj;
    }
    cross[(size_t)i * d + j] += acc;
    if (i == 0) sum[j] += s1;
  }
}

}  // namespace bpm
