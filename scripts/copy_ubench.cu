// What does a persistent tile loop (load a 2^T-amplitude tile into registers, store it) reach on
// B200, and which store mechanism is the fastest?  The fused pass reads and writes 32 B per
// amplitude like a copy, but its memory phase alone runs at 7.5-8 ms per 34 GB where a plain copy
// takes 5.25 ms; stores alone take 5.9 ms (profiles/r02g_oop_phase.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/copy_ubench scripts/copy_ubench.cu
//   scripts/copy_ubench [log2 amplitudes = 30]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::printf("%s: %s\n", #x, cudaGetErrorString(e_));                         \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;

enum { ST_CS = 0, ST_DEFAULT = 1, ST_WT = 2, ST_CG = 3 };

template <int OP>
__device__ __forceinline__ void st128(double2 *p, double a, double b) {
  if (OP == ST_CS) asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  else if (OP == ST_DEFAULT) asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  else if (OP == ST_WT) asm volatile("st.global.wt.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  else asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ld128(const double2 *p, double &a, double &b) {
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}

// MODE bit 0: load, bit 1: store.  Tile = 4096 amplitudes (64 KB), 256 threads x 16 registers pairs;
// thread t, register i <-> element t + 256 i (a warp instruction covers 512 contiguous bytes).
template <int MODE, int OP>
__global__ void __launch_bounds__(256) k_tile(double2 *dst, const double2 *src, u32 ntiles) {
  extern __shared__ unsigned char smem_[];
  double re[16], im[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) re[i] = im[i] = 1e-3 * threadIdx.x;
  for (u32 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const u64 base = (u64)t << 12;
    if (MODE & 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) ld128(src + base + threadIdx.x + 256 * i, re[i], im[i]);
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) st128<OP>(dst + base + threadIdx.x + 256 * i, re[i], im[i]);
    } else {
      double s = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += re[i] + im[i];
      if (s == 1.2345e300) dst[0] = make_double2(s, s);
    }
  }
}

// the same with the NEXT tile's loads issued before the stores of the running one (two register sets)
template <int OP>
__global__ void __launch_bounds__(256) k_tile_pipe(double2 *dst, const double2 *src, u32 ntiles) {
  extern __shared__ unsigned char smem_[];
  double ra[16], ia[16], rb[16], ib[16];
  u32 t = blockIdx.x;
  if (t >= ntiles) return;
#pragma unroll
  for (int i = 0; i < 16; ++i) ld128(src + ((u64)t << 12) + threadIdx.x + 256 * i, ra[i], ia[i]);
  for (;;) {
    const u32 t1 = t + gridDim.x;
    if (t1 < ntiles) {
#pragma unroll
      for (int i = 0; i < 16; ++i) ld128(src + ((u64)t1 << 12) + threadIdx.x + 256 * i, rb[i], ib[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) st128<OP>(dst + ((u64)t << 12) + threadIdx.x + 256 * i, ra[i], ia[i]);
    if (t1 >= ntiles) break;
    const u32 t2 = t1 + gridDim.x;
    if (t2 < ntiles) {
#pragma unroll
      for (int i = 0; i < 16; ++i) ld128(src + ((u64)t2 << 12) + threadIdx.x + 256 * i, ra[i], ia[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) st128<OP>(dst + ((u64)t1 << 12) + threadIdx.x + 256 * i, rb[i], ib[i]);
    if (t2 >= ntiles) break;
    t = t2;
  }
}

// stores through the copy engine: registers -> shared memory -> cp.async.bulk.global.shared::cta
// MODE bit 0: load (LDG into registers), bit 1: store.  NBUF shared buffers of 64 KB.
template <int MODE, int NBUF>
__global__ void __launch_bounds__(256) k_tile_bulkst(double2 *dst, const double2 *src, u32 ntiles) {
  extern __shared__ __align__(128) unsigned char smem_[];
  double re[16], im[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) re[i] = im[i] = 1e-3 * threadIdx.x;
  u32 it = 0;
  for (u32 t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const u64 base = (u64)t << 12;
    if (MODE & 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) ld128(src + base + threadIdx.x + 256 * i, re[i], im[i]);
    }
    unsigned char *buf = smem_ + (size_t)(it % NBUF) * 65536;
    // the bulk store that last read this buffer must have finished READING it
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) *reinterpret_cast<double2 *>(buf + 16 * (threadIdx.x + 256 * i)) = make_double2(re[i], im[i]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const u32 sa = (u32)__cvta_generic_to_shared(buf);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + base), "r"(sa), "r"(65536u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain grid-stride copy, 16 B per thread per iteration (what a library copy looks like)
__global__ void __launch_bounds__(256) k_stream(double2 *dst, const double2 *src, u64 n, int mode) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    double2 v = make_double2(1.0, 2.0);
    if (mode & 1) v = src[i];
    if (mode & 2) dst[i] = v;
    else if (v.x == 1.2345e300) dst[0] = v;
  }
}

template <typename F>
static float timeit(F f, int reps = 3) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char **argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 30;
  const u64 n = 1ull << L;
  const u32 ntiles = (u32)(n >> 12);
  double2 *a = nullptr, *b = nullptr;
  CK(cudaMalloc(&a, n * sizeof(double2)));
  CK(cudaMalloc(&b, n * sizeof(double2)));
  CK(cudaMemset(a, 0, n * sizeof(double2)));
  CK(cudaMemset(b, 0, n * sizeof(double2)));
  const double gb = n * 16.0 / 1e9;
  auto report = [&](const char *name, float ms, double bytes_gb) {
    std::printf("%-58s %7.3f ms  %7.1f GB/s\n", name, ms, bytes_gb / ms * 1e3);
    std::fflush(stdout);
  };
  report("cudaMemcpy D2D (read + write)", timeit([&] { CK(cudaMemcpyAsync(b, a, n * sizeof(double2), cudaMemcpyDeviceToDevice)); }), 2 * gb);
  report("cudaMemset (write only)", timeit([&] { CK(cudaMemsetAsync(b, 0, n * sizeof(double2))); }), gb);
  for (int mode = 1; mode <= 3; ++mode) {
    char nm[96];
    std::snprintf(nm, sizeof nm, "grid-stride 16 B/thread, 148 x 8 CTAs, mode %d (1 ld, 2 st, 3 both)", mode);
    report(nm, timeit([&] { k_stream<<<148 * 8, 256>>>(b, a, n, mode); }), (mode == 3 ? 2 : 1) * gb);
  }
  // tile loop at 1..4 CTAs per SM (occupancy forced by dynamic shared memory)
  for (int occ = 1; occ <= 4; ++occ) {
    const size_t smem = occ == 1 ? 120 * 1024 : (occ == 2 ? 100 * 1024 : (occ == 3 ? 70 * 1024 : 50 * 1024));
#define RUN_TILE(MODE, OP, label)                                                                                   \
  {                                                                                                                 \
    CK(cudaFuncSetAttribute(k_tile<MODE, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    int o = 0;                                                                                                      \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tile<MODE, OP>, 256, smem));                             \
    char nm[96];                                                                                                    \
    std::snprintf(nm, sizeof nm, "tile loop %s, %d CTAs/SM", label, o);                                             \
    report(nm, timeit([&] { k_tile<MODE, OP><<<148 * o, 256, smem>>>(b, a, ntiles); }), (MODE == 3 ? 2 : 1) * gb); \
  }
    RUN_TILE(3, ST_CS, "ld + st.cs")
    RUN_TILE(3, ST_DEFAULT, "ld + st (default)")
    if (occ == 2) {
      RUN_TILE(3, ST_WT, "ld + st.wt")
      RUN_TILE(3, ST_CG, "ld + st.cg")
      RUN_TILE(1, ST_CS, "ld only")
      RUN_TILE(2, ST_CS, "st.cs only")
      RUN_TILE(2, ST_DEFAULT, "st (default) only")
    }
    {
      CK(cudaFuncSetAttribute(k_tile_pipe<ST_CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int o = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tile_pipe<ST_CS>, 256, smem));
      char nm[96];
      std::snprintf(nm, sizeof nm, "tile loop, next loads before the stores, %d CTAs/SM", o);
      report(nm, timeit([&] { k_tile_pipe<ST_CS><<<148 * o, 256, smem>>>(b, a, ntiles); }), 2 * gb);
    }
  }
  // stores through the copy engine
#define RUN_BULK(MODE, NBUF, CTAS, label)                                                                                  \
  {                                                                                                                        \
    const size_t smem = (size_t)NBUF * 65536;                                                                              \
    CK(cudaFuncSetAttribute(k_tile_bulkst<MODE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    int o = 0;                                                                                                             \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tile_bulkst<MODE, NBUF>, 256, smem));                           \
    if (o > CTAS) o = CTAS;                                                                                                \
    char nm[96];                                                                                                           \
    std::snprintf(nm, sizeof nm, "tile loop %s, bulk store from %d smem buffer(s), %d CTAs/SM", label, NBUF, o);           \
    if (o >= 1) report(nm, timeit([&] { k_tile_bulkst<MODE, NBUF><<<148 * o, 256, smem>>>(b, a, ntiles); }), (MODE == 3 ? 2 : 1) * gb); \
  }
  RUN_BULK(3, 1, 1, "ld +")
  RUN_BULK(3, 1, 2, "ld +")
  RUN_BULK(3, 1, 3, "ld +")
  RUN_BULK(3, 2, 1, "ld +")
  RUN_BULK(3, 3, 1, "ld +")
  RUN_BULK(2, 1, 3, "store only,")
  RUN_BULK(2, 3, 1, "store only,")
  return 0;
}
