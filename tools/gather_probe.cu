// Micro-benchmark behind the fused kernel's gather design: how fast can
// one B200 fetch 7 random 800-byte rows per chain (own row + 6 partners) out of an N x 100
// float64 population, as a function of the mechanism and of the bytes kept in flight per SM?
//   ldg   : warp per chain, 14 x LDG.128 per lane issued back to back, W warps per SM
//   bulk  : cp.async.bulk (TMA 1-D) row copies into a shared-memory ring of S chain slots,
//           completion on mbarriers, consumer warps reduce from shared memory
// Each variant sums the 7 rows and writes one 800-byte row per chain (so nothing is elided).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int D = 100, ROWS = 7, ROW_BYTES = D * 8;

__global__ void ldg_kernel(const double* __restrict__ X, const int* __restrict__ idx, int n_chains,
                           double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int c = gw; c < n_chains; c += nw) {
    double2 v[ROWS][2];
    if (lane < 25) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const double* p = X + (size_t)idx[c * ROWS + r] * D + 4 * lane;
        v[r][0] = *reinterpret_cast<const double2*>(p);
        v[r][1] = *reinterpret_cast<const double2*>(p + 2);
      }
      double2 s0 = v[0][0], s1 = v[0][1];
#pragma unroll
      for (int r = 1; r < ROWS; ++r) { s0.x += v[r][0].x; s0.y += v[r][0].y; s1.x += v[r][1].x; s1.y += v[r][1].y; }
      double* o = out + (size_t)c * D + 4 * lane;
      *reinterpret_cast<double2*>(o) = s0;
      *reinterpret_cast<double2*>(o + 2) = s1;
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {      // bounded: a protocol bug must not hang the box
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) return;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ring of S slots, one slot = 7 rows of one chain; warp 0 = producer (lane r issues row r),
// warps 1..CW = consumers, consumer w takes slots w-1, w-1+CW, ...
template <int CW>
__global__ void bulk_kernel(const double* __restrict__ X, const int* __restrict__ idx, int n_chains,
                            double* __restrict__ out, int S) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* ring = reinterpret_cast<double*>(sm);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)S * ROWS * ROW_BYTES);
  uint64_t* empty = full + S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0)
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int n_my = blockIdx.x < n_chains ? (n_chains - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  if (warp == 0) {
    for (int i = 0; i < n_my; ++i) {
      const int s = i % S, round = i / S;
      const int c = blockIdx.x + i * gridDim.x;
      if (round > 0) mbar_wait(empty + s, (round - 1) & 1);
      if (lane == 0) mbar_expect_tx(full + s, ROWS * ROW_BYTES);
      __syncwarp();
      if (lane < ROWS)
        bulk_g2s(ring + ((size_t)s * ROWS + lane) * D, X + (size_t)idx[c * ROWS + lane] * D, ROW_BYTES, full + s);
    }
  } else {
    const int cw = warp - 1;
    for (int i = cw; i < n_my; i += CW) {
      const int s = i % S, round = i / S;
      const int c = blockIdx.x + i * gridDim.x;
      mbar_wait(full + s, round & 1);
      if (lane < 25) {
        const double* base = ring + (size_t)s * ROWS * D + 4 * lane;
        double2 s0 = *reinterpret_cast<const double2*>(base), s1 = *reinterpret_cast<const double2*>(base + 2);
#pragma unroll
        for (int r = 1; r < ROWS; ++r) {
          const double2 a = *reinterpret_cast<const double2*>(base + r * D);
          const double2 b = *reinterpret_cast<const double2*>(base + r * D + 2);
          s0.x += a.x; s0.y += a.y; s1.x += b.x; s1.y += b.y;
        }
        double* o = out + (size_t)c * D + 4 * lane;
        *reinterpret_cast<double2*>(o) = s0;
        *reinterpret_cast<double2*>(o + 2) = s1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  const char* which = argc > 2 ? argv[2] : "all";
  const int n_pop = argc > 1 ? atoi(argv[1]) : 100000;     // population rows
  const int n_chains = n_pop / 2;                           // one half-phase
  double *X, *out; int* idx;
  cudaMalloc(&X, (size_t)n_pop * ROW_BYTES);
  cudaMalloc(&out, (size_t)n_chains * ROW_BYTES);
  cudaMalloc(&idx, sizeof(int) * n_chains * ROWS);
  cudaMemset(X, 0, (size_t)n_pop * ROW_BYTES);
  std::vector<int> h(n_chains * ROWS);
  uint64_t st = 88172645463325252ull;
  for (auto& v : h) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; v = (int)(st % (uint64_t)n_pop); }
  cudaMemcpy(idx, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice);
  double* flush; const size_t FB = 512u << 20; cudaMalloc(&flush, FB);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)n_chains * (ROWS + 1) * ROW_BYTES;
  const int reps = 5;
  printf("population %d rows (%.0f MB), %d chains per launch, %.1f MB per launch\n", n_pop,
         n_pop * 800e-6, n_chains, bytes * 1e-6);
  if (which[0] != 'b')
  for (int wps : {8, 16, 24, 32, 48, 64}) {
    float best = 1e9;
    for (int r = 0; r < reps; ++r) {
      cudaMemsetAsync(flush, r, FB);
      cudaEventRecord(e0);
      ldg_kernel<<<148 * (wps / 8), 256>>>(X, idx, n_chains, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      best = fminf(best, time_ms(e0, e1));
    }
    printf("ldg  warps/SM %2d : %7.1f us  %7.1f GB/s\n", wps, best * 1e3, bytes / best * 1e-6);
  }
  if (which[0] != 'l')
  for (int S : {8, 16, 24, 32}) {   // multiples of CW: a slot is always drained by the same warp
    const size_t smem = (size_t)S * ROWS * ROW_BYTES + 16 * S + 128;
    cudaFuncSetAttribute(bulk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e9;
    for (int r = 0; r < reps; ++r) {
      cudaMemsetAsync(flush, r, FB);
      cudaEventRecord(e0);
      bulk_kernel<8><<<148, 32 * 9, smem>>>(X, idx, n_chains, out, S);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      best = fminf(best, time_ms(e0, e1));
    }
    printf("bulk slots/SM %2d (%5.1f KB in flight) : %7.1f us  %7.1f GB/s   [%s]\n", S, S * ROWS * ROW_BYTES / 1024.0,
           best * 1e3, bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
