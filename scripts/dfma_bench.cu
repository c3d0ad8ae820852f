// microbenchmark: DFMA throughput with a vector-register vs a uniform-register multiplier
#include <cstdio>
#include <cuda_runtime.h>
struct P { double m[8]; };
template <bool UNI>
__global__ void __launch_bounds__(256) k(double *out, const __grid_constant__ P p, int iters) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
  double a = p.m[0], b = p.m[1];
  if (!UNI) { a += threadIdx.x * 1e-30; b += threadIdx.x * 1e-30; }  // per-thread values -> vector registers
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(a, x[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main() {
  double *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  P p; for (int i = 0; i < 8; ++i) p.m[i] = 0.999 + i * 1e-4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 1; warps <= 2; ++warps)
  for (int uni = 0; uni < 2; ++uni) {
    const int grid = 148 * (warps == 1 ? 2 : 4);  // 16 or 32 warps per SM
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (uni) k<true><<<grid, 256>>>(out, p, iters); else k<false><<<grid, 256>>>(out, p, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma = double(grid) * 256 * 16 * iters;
    printf("uniform=%d warps/SM=%d: %.3f ms, %.2f TDFMA/s, %.1f DFMA/clk/SM @1.965GHz\n", uni, grid * 8 / 148, ms, dfma / ms / 1e9, dfma / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
