// Probe: peak FP64 rate on B200 through (a) plain DFMA, (b) mma.sync m8n8k4 f64 (DMMA),
// (c) mma.sync m16n8k16 f64.  Decides how the quadratic-form tile is computed.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double x = 1.000001, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void dmma884_kernel(double* out, int iters) {
  double c[4][2];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

__global__ void dmma16816_kernel(double* out, int iters) {
  double c[2][4];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  double* out;
  cudaMalloc(&out, 8);
  const int blocks = 148 * 8, threads = 256, iters = 20000;
  float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
  double fl = 2.0 * 8 * iters * (double)blocks * threads;
  printf("DFMA      : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
  ms = time_it([&] { dmma884_kernel<<<blocks, threads>>>(out, iters); });
  fl = 2.0 * 8 * 8 * 4 * 4.0 * iters * (double)blocks * (threads / 32);
  printf("DMMA 8x8x4: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
  ms = time_it([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters); });
  fl = 2.0 * 16 * 8 * 16 * 2.0 * iters * (double)blocks * (threads / 32);
  printf("DMMA 16x8x16: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
