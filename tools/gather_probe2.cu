// Fair version of tools/gather_probe.cu: can TMA row copies feed the 7-row
// gather of the fused kernel (own row + 6 partners, 800 B each, out of an N x 100 f64 population)?
// The round-1 probe had ONE producer warp per SM and a dependent global index load per chain, so it
// measured a serial latency loop (~0.8 us per chain per SM), not the copy engine.  Here
//   * partner indices come from a hash in registers (the real kernel's are Philox outputs),
//   * EVERY warp issues: a warp owns K ring slots of one chain each (7 rows = 5.6 KB), keeps K chains in
//     flight, consumes a slot from shared memory (sums the rows, writes one row) and refills it,
//   * three mechanisms: ldg (landing registers, the shipped design), bulk (cp.async.bulk 1-D rows,
//     UBLKCP), gather4 (cp.async.bulk.tensor.2d tile::gather4: 4 arbitrary rows per instruction, UTMALDG).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe2 tools/gather_probe2.cu
// Run:   tools/gather_probe2 [n_pop] [ldg|bulk|g4|all]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

constexpr int D = 100, ROWS = 7, ROW_BYTES = D * 8;
constexpr int SLOT_ROWS = 8;                       // gather4 lands 2 x 4 rows; bulk uses 7 of the 8
constexpr int SLOT_BYTES = SLOT_ROWS * ROW_BYTES;  // 6400

__host__ __device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ int row_of(int c, int r, int n_pop) {
  return (int)(((uint64_t)mix((uint32_t)c * 8u + (uint32_t)r + 0x9E3779B9u) * (uint64_t)n_pop) >> 32);
}

__global__ void ldg_kernel(const double* __restrict__ X, int n_pop, int n_chains, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int c = gw; c < n_chains; c += nw) {
    double2 v[ROWS][2];
    if (lane < 25) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const double* p = X + (size_t)row_of(c, r, n_pop) * D + 4 * lane;
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
__device__ __forceinline__ void gather4_g2s(void* dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
               "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
               : "memory");
}

// Every warp: K slots, chains c = gw, gw + nw, ...; MODE 0 = cp.async.bulk rows, 1 = gather4 (2 instructions per chain)
template <int MODE>
__global__ void ring_kernel(const double* __restrict__ X, const __grid_constant__ CUtensorMap tm, int n_pop,
                            int n_chains, double* __restrict__ out, int K) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  unsigned char* ring = sm + (size_t)warp * K * SLOT_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)W * K * SLOT_BYTES) + warp * K;
  if (lane == 0)
    for (int s = 0; s < K; ++s) mbar_init(full + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int gw = blockIdx.x * W + warp, nw = gridDim.x * W;
  const int n_my = gw < n_chains ? (n_chains - 1 - gw) / nw + 1 : 0;
  auto issue = [&](int i) {
    const int s = i % K, c = gw + i * nw;
    double* dst = reinterpret_cast<double*>(ring + (size_t)s * SLOT_BYTES);
    if (MODE == 0) {
      if (lane == 0) mbar_expect_tx(full + s, ROWS * ROW_BYTES);
      __syncwarp();
      if (lane < ROWS) bulk_g2s(dst + lane * D, X + (size_t)row_of(c, lane, n_pop) * D, ROW_BYTES, full + s);
    } else {
      if (lane == 0) {
        mbar_expect_tx(full + s, 8 * ROW_BYTES);
        gather4_g2s(dst, &tm, 0, row_of(c, 0, n_pop), row_of(c, 1, n_pop), row_of(c, 2, n_pop), row_of(c, 3, n_pop),
                    full + s);
        gather4_g2s(dst + 4 * D, &tm, 0, row_of(c, 4, n_pop), row_of(c, 5, n_pop), row_of(c, 6, n_pop),
                    row_of(c, 6, n_pop), full + s);
      }
    }
  };
  for (int i = 0; i < K && i < n_my; ++i) issue(i);
  for (int i = 0; i < n_my; ++i) {
    const int s = i % K, c = gw + i * nw;
    mbar_wait(full + s, (i / K) & 1);
    if (lane < 25) {
      const double* base = reinterpret_cast<const double*>(ring + (size_t)s * SLOT_BYTES) + 4 * lane;
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads before the async refill
    if (i + K < n_my) issue(i + K);
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  const int n_pop = argc > 1 ? atoi(argv[1]) : 100000;
  const char* which = argc > 2 ? argv[2] : "all";
  const int n_chains = n_pop / 2;
  double *X, *out;
  cudaMalloc(&X, (size_t)n_pop * ROW_BYTES);
  cudaMalloc(&out, (size_t)n_chains * ROW_BYTES);
  cudaMemset(X, 0, (size_t)n_pop * ROW_BYTES);
  double* flush; const size_t FB = 512u << 20; cudaMalloc(&flush, FB);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)n_chains * (ROWS + 1) * ROW_BYTES;
  const int reps = 5;
  printf("population %d rows (%.0f MB), %d chains per launch, %.1f MB per launch, indices from registers\n", n_pop,
         n_pop * 800e-6, n_chains, bytes * 1e-6);
  const bool all = !strcmp(which, "all");
  if (all || !strcmp(which, "ldg"))
    for (int wps : {8, 16, 24, 32}) {
      float best = 1e9;
      for (int r = 0; r < reps; ++r) {
        cudaMemsetAsync(flush, r, FB);
        cudaEventRecord(e0);
        ldg_kernel<<<148 * (wps / 8), 256>>>(X, n_pop, n_chains, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        best = fminf(best, time_ms(e0, e1));
      }
      printf("ldg     warps/SM %2d                              : %7.1f us  %7.1f GB/s\n", wps, best * 1e3,
             bytes / best * 1e-6);
    }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  bool have_tm = false;
  if (all || !strcmp(which, "g4")) {
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) == cudaSuccess && enc) {
      cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)n_pop};
      cuuint64_t gstr[1] = {(cuuint64_t)ROW_BYTES};
      cuuint32_t box[2] = {(cuuint32_t)D, 1};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      have_tm = r == CUDA_SUCCESS;
      if (!have_tm) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    } else {
      printf("cuTensorMapEncodeTiled entry point unavailable\n");
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 0 && !(all || !strcmp(which, "bulk"))) continue;
    if (mode == 1 && !have_tm) continue;
    const int cfgs[][2] = {{4, 2}, {4, 8}, {8, 2}, {8, 4}, {16, 1}, {16, 2}, {32, 1}};
    for (auto& wc : cfgs) {
      const int W = wc[0], K = wc[1];
      const size_t smem = (size_t)W * K * SLOT_BYTES + 8 * W * K + 128;
      if (smem > 227 * 1024) continue;
      if (mode == 0) cudaFuncSetAttribute(ring_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      else cudaFuncSetAttribute(ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float best = 1e9;
      for (int r = 0; r < reps; ++r) {
        cudaMemsetAsync(flush, r, FB);
        cudaEventRecord(e0);
        if (mode == 0) ring_kernel<0><<<148, 32 * W, smem>>>(X, tm, n_pop, n_chains, out, K);
        else ring_kernel<1><<<148, 32 * W, smem>>>(X, tm, n_pop, n_chains, out, K);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        best = fminf(best, time_ms(e0, e1));
      }
      cudaError_t err = cudaGetLastError();
      printf("%-7s warps/SM %2d x %d chains in flight (%5.1f KB) : %7.1f us  %7.1f GB/s   [%s]\n",
             mode == 0 ? "bulk" : "gather4", W, K, W * K * ROWS * ROW_BYTES / 1024.0, best * 1e3, bytes / best * 1e-6,
             cudaGetErrorString(err));
      if (err != cudaSuccess) return 1;
    }
  }
  return 0;
}
