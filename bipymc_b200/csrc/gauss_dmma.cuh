// Batched Gaussian quadratic form for LARGE d on the FP64 tensor pipe:
//   out[i] = finish(c0, | (P[i] - mu) . W |^2),   P [n][ld], W [d][r] row-major
// (scipy _logpdf's maha = sum(square(dot(dev, prec_U)), axis=-1), _multivariate.py:566-591; the
// reference's 1000-D target, d100_gauss.py:14-35 with dim = 1000 -- SURVEY.md section 8a L3, the
// only dense contraction on the path).  At d = 1000 a chain-step is 2 d r = 2e6 FP64 flop against
// ~100 kB of traffic: FP64-bound (20 flop/B vs a ridge of ~6), so this is a GEMM kernel.
//   * CTA = 8 warps, output tile 64 rows x 128 columns, warp tile 32 x 32 = 4 x 4 DMMA m8n8k4
//     tiles (16 independent accumulator pairs per warp, 8 fragment loads per 16 MMAs);
//   * k tiles of 16 staged in shared memory through a 3-deep cp.async ring (A 64 x 16, B 16 x 128),
//     row strides 20 / 132 doubles so every fragment load is bank-conflict free;
//   * Y never reaches memory: each column block's tile is squared and folded into per-row sums,
//     reduced over the quad lanes and over the 4 warps that share a row group at the end.
// tcgen05 has no f64 MMA kind; mma.sync.m8n8k4.f64 is the FP64 tensor instruction of sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "targets.cuh"

namespace bpm {

constexpr int kGdN = 128, kGdK = 16, kGdStages = 3;     // CTA tile: (16 MT) rows x 128 columns
constexpr int kGdLda = kGdK + 4;      // 20: (4 row + k) mod 16 distinct over a half warp
constexpr int kGdLdb = kGdN + 4;      // 132
constexpr int kGdThreads = 256;
__host__ __device__ constexpr size_t gd_stage_doubles(int M) { return (size_t)M * kGdLda + (size_t)kGdK * kGdLdb; }

__device__ __forceinline__ void gd_cp16(void* dst, const void* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = ok ? 16 : 0;                      // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void gd_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// requires d % 2 == 0, r % 2 == 0, ld % 2 == 0 (16-byte chunks).  MT = DMMA m-tiles per warp: 4 -> 64-row
// CTAs (best reuse), 2 -> 32-row CTAs for launches too small to give every SM two 64-row CTAs.
template <bool CENTER, int MT>
__global__ void __launch_bounds__(kGdThreads, 2)
gauss_dmma_kernel(const double* __restrict__ P, int n, int ld, int d, int r, const double* __restrict__ mu,
                  const double* __restrict__ W, double c0, int log_of_pdf, double* __restrict__ out,
                  const int32_t* __restrict__ n_dev) {
  constexpr int kGdM = 16 * MT;
  if (n_dev) {                       // rows actually present (device-counted packed list); n is the launch bound
    n = min(n, *n_dev);
    if ((int)blockIdx.x * kGdM >= n) return;
  }
  constexpr size_t kGdStageDoubles = gd_stage_doubles(kGdM);
  extern __shared__ __align__(16) double gsm[];
  double* mus = gsm;                                           // [dpad]
  const int dpad = (d + kGdK - 1) / kGdK * kGdK;
  double* ring = gsm + dpad;
  double* red = ring + kGdStages * kGdStageDoubles;            // [4][64] row sums per column-warp
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 2, wc = warp & 3;                     // warp tile: rows 8 MT wr, cols 32 wc
  const int lq = lane >> 2, lk = lane & 3;
  const int m0 = blockIdx.x * kGdM;
  if (CENTER)
    for (int k = tid; k < dpad; k += kGdThreads) mus[k] = k < d ? mu[k] : 0.0;
  const int nkt = dpad / kGdK;
  const int ncb = (r + kGdN - 1) / kGdN;
  const int total = ncb * nkt;                                 // pipeline steps: column block major

  auto issue = [&](int step) {
    if (step < total) {
      const int cb = step / nkt, kt = step - cb * nkt;
      double* As = ring + (size_t)(step % kGdStages) * kGdStageDoubles;
      double* Bs = As + (size_t)kGdM * kGdLda;
      // A: 64 rows x 16 doubles = 512 chunks of 16 B
      for (int c = tid; c < kGdM * (kGdK / 2); c += kGdThreads) {
        const int row = c >> 3, k2 = (c & 7) * 2;
        const int gr = m0 + row, gk = kt * kGdK + k2;
        const bool ok = gr < n && gk < d;
        gd_cp16(As + row * kGdLda + k2, P + (size_t)(ok ? gr : 0) * ld + (ok ? gk : 0), ok);
      }
      // B: 16 rows x 128 doubles = 1024 chunks
      for (int c = tid; c < kGdK * (kGdN / 2); c += kGdThreads) {
        const int kk = c >> 6, n2 = (c & 63) * 2;
        const int gk = kt * kGdK + kk, gn = cb * kGdN + n2;
        const bool ok = gk < d && gn < r;
        gd_cp16(Bs + kk * kGdLdb + n2, W + (size_t)(ok ? gk : 0) * r + (ok ? gn : 0), ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue(0);
  issue(1);
  double rs[MT][1];                                            // per m-tile partial row sums of this lane
#pragma unroll
  for (int i = 0; i < MT; ++i) rs[i][0] = 0.0;
  double acc[MT][4][2];
  for (int step = 0; step < total; ++step) {
    const int cb = step / nkt, kt = step - cb * nkt;
    if (kt == 0) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");       // this step's tiles have landed
    __syncthreads();                                           // ... for every thread; the slot of step-1 is free
    issue(step + 2);
    const double* As = ring + (size_t)(step % kGdStages) * kGdStageDoubles;
    const double* Bs = As + (size_t)kGdM * kGdLda;
#pragma unroll
    for (int k4 = 0; k4 < kGdK / 4; ++k4) {
      double af[MT], bf[4];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        af[i] = As[(8 * MT * wr + 8 * i + lq) * kGdLda + 4 * k4 + lk];
        if (CENTER) af[i] = __dsub_rn(af[i], mus[kt * kGdK + 4 * k4 + lk]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(4 * k4 + lk) * kGdLdb + 32 * wc + 8 * j + lq];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gd_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    if (kt == nkt - 1) {                                       // column block finished: fold its squares
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rs[i][0] = fma(acc[i][j][0], acc[i][j][0], rs[i][0]);
          rs[i][0] = fma(acc[i][j][1], acc[i][j][1], rs[i][0]);
        }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // quad lanes hold disjoint columns of the same rows; then the 4 column-warps of a row group
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    double v = rs[i][0];
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
    if (lk == 0) red[wc * kGdM + 8 * MT * wr + 8 * i + lq] = v;
  }
  __syncthreads();
  if (tid < kGdM) {
    const int row = m0 + tid;
    if (row < n) {
      double maha = red[tid];
      maha = __dadd_rn(maha, red[kGdM + tid]);
      maha = __dadd_rn(maha, red[2 * kGdM + tid]);
      maha = __dadd_rn(maha, red[3 * kGdM + tid]);
      out[row] = gauss_finish(c0, maha, log_of_pdf);
    }
  }
}

inline size_t gauss_dmma_smem(int d, int M) {
  const int dpad = (d + kGdK - 1) / kGdK * kGdK;
  return sizeof(double) * ((size_t)dpad + kGdStages * gd_stage_doubles(M) + 4 * (size_t)M);
}
inline bool gauss_dmma_supported(int d, int r, int ld) { return (d % 2) == 0 && (r % 2) == 0 && (ld % 2) == 0 && d >= 16; }

template <bool CENTER, int MT>
inline int launch_gauss_dmma_t(const double* P, int n, int ld, int d, int r, const double* mu, const double* W,
                               double c0, int log_of_pdf, double* out, const int32_t* n_dev, cudaStream_t s) {
  constexpr int M = 16 * MT;
  const size_t sm = gauss_dmma_smem(d, M);
  cudaError_t e = cudaFuncSetAttribute(gauss_dmma_kernel<CENTER, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return 1;
  gauss_dmma_kernel<CENTER, MT><<<(n + M - 1) / M, kGdThreads, sm, s>>>(P, n, ld, d, r, mu, W, c0, log_of_pdf, out,
                                                                      n_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

inline int launch_gauss_dmma(const double* P, int n, int ld, int d, int r, const double* mu, const double* W,
                             double c0, int log_of_pdf, int mu_is_zero, double* out, const int32_t* n_dev,
                             cudaStream_t s) {
  // 64-row CTAs once there are enough of them to put two on every SM, else 32-row CTAs
  const bool small = (n + 63) / 64 < 2 * 148;
  if (mu_is_zero)
    return small ? launch_gauss_dmma_t<false, 2>(P, n, ld, d, r, mu, W, c0, log_of_pdf, out, n_dev, s)
                 : launch_gauss_dmma_t<false, 4>(P, n, ld, d, r, mu, W, c0, log_of_pdf, out, n_dev, s);
  return small ? launch_gauss_dmma_t<true, 2>(P, n, ld, d, r, mu, W, c0, log_of_pdf, out, n_dev, s)
               : launch_gauss_dmma_t<true, 4>(P, n, ld, d, r, mu, W, c0, log_of_pdf, out, n_dev, s);
}

}  // namespace bpm
