// Microbenchmark: cost of register-resident 1-qubit gate bodies behind an interpreter-style
// switch (no memory traffic), 16 complex amplitudes per thread as in k_fused_pass<12,4>.
//   M0 real 2x2, 4 ops per component pair (the shipped gate_real)
//   M1 rotation as three in-place shears (3 DFMA per component pair, no temporaries)
//   M2 scaled rotation [[1,p],[q,1]] (2 DFMA per component pair, needs one temporary)
//   M3 general complex 2x2 (8 ops per amplitude, the shipped gate_general)
//   M4 general complex scaled so that m00 = 1 (6 ops per amplitude)
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/gate_ubench.cu -o scripts/gate_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Prog {
  int n;
  int op[64];
  double c[64][8];
};

template <int M, int J>
__device__ __forceinline__ void gate(double (&re)[16], double (&im)[16], const double *c) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int i0 = ((p >> J) << (J + 1)) | (p & ((1 << J) - 1));
    const int i1 = i0 | (1 << J);
    if (M == 0) {
      const double a = c[0], b = c[1], cc = c[2], d = c[3];
      const double Tr = cc * re[i0], Ti = cc * im[i0], Pr = b * re[i1], Pi = b * im[i1];
      re[i0] = fma(a, re[i0], Pr);
      im[i0] = fma(a, im[i0], Pi);
      re[i1] = fma(d, re[i1], Tr);
      im[i1] = fma(d, im[i1], Ti);
    } else if (M == 1) {
      const double t = c[0], s = c[1];
      re[i0] = fma(t, re[i1], re[i0]);
      im[i0] = fma(t, im[i1], im[i0]);
      re[i1] = fma(s, re[i0], re[i1]);
      im[i1] = fma(s, im[i0], im[i1]);
      re[i0] = fma(t, re[i1], re[i0]);
      im[i0] = fma(t, im[i1], im[i0]);
    } else if (M == 2) {
      const double pp = c[0], q = c[1];
      const double yr = fma(pp, re[i1], re[i0]), yi = fma(pp, im[i1], im[i0]);
      re[i1] = fma(q, re[i0], re[i1]);
      im[i1] = fma(q, im[i0], im[i1]);
      re[i0] = yr;
      im[i0] = yi;
    } else if (M == 3) {
      const double Ar = c[0], Ai = c[1], Br = c[2], Bi = c[3], Cr = c[4], Ci = c[5], Dr = c[6], Di = c[7];
      double P = -Ai * im[i0], Q = Ai * re[i0], Tr = Cr * re[i0], Ti = Cr * im[i0];
      P = fma(Br, re[i1], P);
      Q = fma(Br, im[i1], Q);
      Tr = fma(-Ci, im[i0], Tr);
      Ti = fma(Ci, re[i0], Ti);
      P = fma(-Bi, im[i1], P);
      Q = fma(Bi, re[i1], Q);
      Tr = fma(-Di, im[i1], Tr);
      Ti = fma(Di, re[i1], Ti);
      re[i0] = fma(Ar, re[i0], P);
      im[i0] = fma(Ar, im[i0], Q);
      re[i1] = fma(Dr, re[i1], Tr);
      im[i1] = fma(Dr, im[i1], Ti);
    } else if (M == 4) {
      const double Br = c[2], Bi = c[3], Cr = c[4], Ci = c[5], Dr = c[6], Di = c[7];
      double Tr = Cr * re[i0], Ti = Cr * im[i0];
      double P = fma(Br, re[i1], re[i0]), Q = fma(Br, im[i1], im[i0]);
      Tr = fma(-Ci, im[i0], Tr);
      Ti = fma(Ci, re[i0], Ti);
      Tr = fma(-Di, im[i1], Tr);
      Ti = fma(Di, re[i1], Ti);
      re[i0] = fma(-Bi, im[i1], P);
      im[i0] = fma(Bi, re[i1], Q);
      re[i1] = fma(Dr, re[i1], Tr);
      im[i1] = fma(Dr, im[i1], Ti);
    }
  }
}

template <int M, int MINB>
__global__ void __launch_bounds__(256, MINB) k(double2 *out, const __grid_constant__ Prog prog, int iters) {
  double re[16], im[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    re[i] = threadIdx.x * 1e-3 + i;
    im[i] = threadIdx.x * 2e-3 - i;
  }
  for (int it = 0; it < iters; ++it) {
    for (int g = 0; g < prog.n; ++g) {
      const double *c = prog.c[g];
      switch (prog.op[g]) {
        case 0: gate<M, 0>(re, im, c); break;
        case 1: gate<M, 1>(re, im, c); break;
        case 2: gate<M, 2>(re, im, c); break;
        case 3: gate<M, 3>(re, im, c); break;
        default: break;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) out[(size_t(blockIdx.x) * 256 + threadIdx.x) * 16 + i] = make_double2(re[i], im[i]);
}

template <int M, int MINB>
void run(const char *name, int fp64_per_gate, double2 *out, const Prog &p, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * MINB;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    k<M, MINB><<<grid, 256>>>(out, p, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, (const void *)k<M, MINB>);
  const double gates = double(iters) * p.n;             // per thread
  const double warps_per_smsp = MINB * 8 / 4.0;
  const double cyc = ms * 1e-3 * 1.93e9 / (gates * warps_per_smsp);  // SMSP cycles per warp-gate
  // time to apply one such gate to 2^30 amplitudes at this rate (16 amplitudes per thread)
  const double ms30 = ms / gates / (double(grid) * 256 * 16) * 1073741824.0;
  printf("%-28s ctas/SM=%d regs=%3d: %8.3f ms  %6.1f cyc/warp-gate (fp64-pipe floor %d)  -> %.3f ms per gate at 30q\n", name,
         MINB, fa.numRegs, ms, cyc, 2 * fp64_per_gate, ms30);
}

int main() {
  double2 *out;
  cudaMalloc(&out, sizeof(double2) * 148 * 4 * 256 * 16);
  Prog p{};
  p.n = 52;
  for (int g = 0; g < p.n; ++g) {
    p.op[g] = g % 4;
    const double th = 0.3 + 0.01 * g;
    // M0: rotation; M1: shears (t, s); M2: tiny p, q so that nothing overflows; M3/M4: complex
    p.c[g][0] = 0.0;
  }
  Prog p0 = p, p1 = p, p2 = p, p3 = p;
  for (int g = 0; g < p.n; ++g) {
    const double th = 0.3 + 0.01 * g;
    p0.c[g][0] = cos(th); p0.c[g][1] = -sin(th); p0.c[g][2] = sin(th); p0.c[g][3] = cos(th);
    p1.c[g][0] = -tan(th / 2); p1.c[g][1] = sin(th);
    p2.c[g][0] = -1e-9 * (g + 1); p2.c[g][1] = 1e-9 * (g + 1);
    const double u = 1 / sqrt(2.0);
    p3.c[g][0] = cos(th) * u; p3.c[g][1] = cos(th) * u; p3.c[g][2] = -sin(th) * u; p3.c[g][3] = sin(th) * u;
    p3.c[g][4] = sin(th) * u; p3.c[g][5] = sin(th) * u; p3.c[g][6] = cos(th) * u; p3.c[g][7] = -cos(th) * u;
  }
  const int iters = 2000;
  run<0, 3>("M0 real 4-op", 64, out, p0, iters);
  run<0, 2>("M0 real 4-op", 64, out, p0, iters);
  run<1, 3>("M1 three shears", 48, out, p1, iters);
  run<1, 2>("M1 three shears", 48, out, p1, iters);
  run<1, 4>("M1 three shears", 48, out, p1, iters);
  run<2, 3>("M2 scaled rotation 2-op", 32, out, p2, iters);
  run<2, 2>("M2 scaled rotation 2-op", 32, out, p2, iters);
  run<2, 4>("M2 scaled rotation 2-op", 32, out, p2, iters);
  run<3, 3>("M3 general 8-op", 128, out, p3, iters);
  run<3, 2>("M3 general 8-op", 128, out, p3, iters);
  run<4, 3>("M4 general scaled 6-op", 96, out, p3, iters);
  run<4, 2>("M4 general scaled 6-op", 96, out, p3, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
