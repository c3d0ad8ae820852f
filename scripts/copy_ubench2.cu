// Memory floor of the fused pass's tile loop on B200, by access pattern, shared-memory footprint
// (= what is left of the 228 KB for L1, where in-flight global loads park their lines) and CTAs
// per SM.  Tile = 4096 amplitudes selected by 12 physical index bits; the other bits enumerate tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/copy_ubench2 scripts/copy_ubench2.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                \
  do {                                                       \
    cudaError_t e_ = (x);                                    \
    if (e_ != cudaSuccess) {                                 \
      std::printf("%s: %s\n", #x, cudaGetErrorString(e_));   \
      std::exit(1);                                          \
    }                                                        \
  } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;

struct Geo {
  unsigned char in_pos[12], out_pos[12];  // physical bit of tile-local bit i (load / store side)
  u32 in_nruns, out_nruns;
  u32 in_shift[8], in_len[8], out_shift[8], out_len[8];
  int prefetch;  // L2 prefetch of the CTA's next tile: 1 per 128-byte line, 2 bulk per chunk, 3 bulk + evict_first hint,
                 // 4 per line with L2::evict_last, 5 per line but only AFTER the delay (late)
  int mode;      // bit 0 load, bit 1 store
  int delay;     // SM cycles each CTA idles per tile between its loads and its stores (stands for the gates)
  int chunk_bits;  // contiguous low tile bits
};

__device__ __forceinline__ u64 deposit(u32 t, u32 nruns, const u32 *shift, const u32 *len) {
  u64 b = 0;
  u64 x = t;
  for (u32 k = 0; k < nruns; ++k) {
    b |= (x & ((1ull << len[k]) - 1ull)) << shift[k];
    x >>= len[k];
  }
  return b;
}

// thread bits 0..7 <-> tile-local bits 0..7, register bits <-> tile-local bits 8..11
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_tile(double2 *dst, const double2 *src, u32 ntiles, const __grid_constant__ Geo g) {
  extern __shared__ unsigned char smem_[];
  u64 *tab = reinterpret_cast<u64 *>(smem_);
  const u32 tid = threadIdx.x;
  u64 oi = 0, oo = 0;
  for (int j = 0; j < 8; ++j) {
    oi |= (u64)((tid >> j) & 1u) << g.in_pos[j];
    oo |= (u64)((tid >> j) & 1u) << g.out_pos[j];
  }
  u64 ri[16], ro[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    ri[i] = ro[i] = 0;
    for (int j = 0; j < 4; ++j) {
      ri[i] |= (u64)((i >> j) & 1) << g.in_pos[8 + j];
      ro[i] |= (u64)((i >> j) & 1) << g.out_pos[8 + j];
    }
  }
  // the 128-byte lines of a tile this thread prefetches (2 of 512)
  u64 pl[2];
  for (int k = 0; k < 2; ++k) {
    const u32 l = tid + 256 * k;
    pl[k] = 0;
    for (int j = 0; j < 9; ++j) pl[k] |= (u64)((l >> j) & 1u) << g.in_pos[3 + j];
  }
  (void)tab;
  double re[16], im[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) re[i] = im[i] = 1e-3 * tid;
  for (u32 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const u64 bi = deposit(t, g.in_nruns, g.in_shift, g.in_len);
    if (g.mode & 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(re[i]), "=d"(im[i]) : "l"(src + bi + oi + ri[i]));
      if (g.prefetch && g.prefetch != 5 && t + gridDim.x < ntiles) {
        const u64 nb = deposit(t + gridDim.x, g.in_nruns, g.in_shift, g.in_len);
        if (g.prefetch == 1) {
          for (int k = 0; k < 2; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + nb + pl[k]));
        } else if (g.prefetch == 4) {
          for (int k = 0; k < 2; ++k) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(src + nb + pl[k]));
        } else {
          const u32 nchunks = 1u << (12 - g.chunk_bits);
          if (tid < nchunks) {
            u64 co = 0;
            for (int j = 0; j < 12 - g.chunk_bits; ++j) co |= (u64)((tid >> j) & 1u) << g.in_pos[g.chunk_bits + j];
            const u32 bytes = 16u << g.chunk_bits;
            if (g.prefetch == 2) {
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + nb + co), "r"(bytes) : "memory");
            } else {
              u64 pol;
              asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
              asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src + nb + co), "r"(bytes), "l"(pol) : "memory");
            }
          }
        }
      }
    }
    if (g.delay > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < g.delay) __nanosleep(200);
    }
    if ((g.mode & 1) && g.prefetch == 5 && t + gridDim.x < ntiles) {
      const u64 nb = deposit(t + gridDim.x, g.in_nruns, g.in_shift, g.in_len);
      for (int k = 0; k < 2; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + nb + pl[k]));
    }
    if (g.mode & 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) re[i] *= 1.0000001;  // (the data really changes)
      const u64 bo = deposit(t, g.out_nruns, g.out_shift, g.out_len);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(dst + bo + oo + ro[i]), "d"(re[i]), "d"(im[i]) : "memory");
    } else {
      double s = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += re[i] + im[i];
      if (s == 1.2345e300) dst[0] = make_double2(s, s);
    }
  }
}

__global__ void k_fill(double2 *p, u64 n) {
  for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    u64 x = i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 32;
    p[i] = make_double2((double)(x & 0xffffff) * 1e-7 + 0.1, (double)((x >> 24) & 0xffffff) * 1e-7 - 0.7);
  }
}

static void runs_of(const unsigned char *pos, int L, u32 &nruns, u32 *shift, u32 *len) {
  u64 mask = 0;
  for (int i = 0; i < 12; ++i) mask |= 1ull << pos[i];
  nruns = 0;
  int b = 0;
  while (b < L) {
    if (mask & (1ull << b)) {
      ++b;
      continue;
    }
    int e = b;
    while (e < L && !(mask & (1ull << e))) ++e;
    shift[nruns] = b;
    len[nruns] = e - b;
    ++nruns;
    b = e;
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
  k_fill<<<148 * 8, 256>>>(a, n);  // (incompressible data: constant fills read above the HBM peak)
  k_fill<<<148 * 8, 256>>>(b, n);
  CK(cudaDeviceSynchronize());
  const double gb = n * 16.0 / 1e9;
  struct Pat {
    const char *name;
    unsigned char pos[12];
  };
  const Pat pats[] = {
      {"contiguous 64 KB", {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11}},
      {"128 B chunks (benchmark-like)", {0, 1, 2, 9, 12, 14, 17, 18, 21, 23, 26, 28}},
      {"512 B chunks", {0, 1, 2, 3, 4, 12, 14, 17, 18, 21, 23, 26}},
      {"2 KB chunks", {0, 1, 2, 3, 4, 5, 6, 14, 17, 21, 23, 26}},
      {"pass-0-like: 128 B chunks on 512 pages", {0, 1, 2, 21, 22, 23, 24, 25, 26, 27, 28, 29}},
      {"pass-6-like: 128 B chunks on 256 pages", {0, 1, 2, 10, 17, 18, 22, 23, 24, 25, 28, 29}},
      {"128 B chunks inside 2 MB pages", {0, 1, 2, 5, 6, 8, 9, 10, 12, 13, 15, 16}},
      {"128 B chunks on 8 pages", {0, 1, 2, 5, 6, 8, 9, 10, 12, 17, 22, 27}},
      {"128 B chunks on 64 pages", {0, 1, 2, 5, 6, 8, 17, 19, 22, 24, 27, 29}},
      {"re-sorted layout: 512 B chunks, bits 0-4 + 12-18", {0, 1, 2, 3, 4, 12, 13, 14, 15, 16, 17, 18}},
      {"re-sorted layout, low_bits 3: bits 0-2 + 12-20", {0, 1, 2, 12, 13, 14, 15, 16, 17, 18, 19, 20}},
  };
  int delay = 0;
  // explicit tile-number deposits (load side / store side): lists of (len, shift), empty = ascending runs
  std::vector<std::pair<int, int>> in_runs, out_runs;
  auto run = [&](const char *label, const unsigned char *ip, const unsigned char *op, bool inplace, int mode, int prefetch, int ctas,
                 size_t smem) {
    Geo g{};
    g.delay = delay;
    g.chunk_bits = 3;
    while (g.chunk_bits < 12 && ip[g.chunk_bits] == g.chunk_bits) ++g.chunk_bits;
    std::memcpy(g.in_pos, ip, 12);
    std::memcpy(g.out_pos, op, 12);
    runs_of(ip, L, g.in_nruns, g.in_shift, g.in_len);
    runs_of(op, L, g.out_nruns, g.out_shift, g.out_len);
    if (!in_runs.empty()) {
      g.in_nruns = (u32)in_runs.size();
      for (size_t k = 0; k < in_runs.size(); ++k) g.in_len[k] = in_runs[k].first, g.in_shift[k] = in_runs[k].second;
    }
    if (!out_runs.empty()) {
      g.out_nruns = (u32)out_runs.size();
      for (size_t k = 0; k < out_runs.size(); ++k) g.out_len[k] = out_runs[k].first, g.out_shift[k] = out_runs[k].second;
    }
    g.prefetch = prefetch;
    g.mode = mode;
    auto launch = [&](auto kern) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int o = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, 256, smem));
      if (o > ctas) o = ctas;
      const float ms = timeit([&] { kern<<<148 * o, 256, smem>>>(inplace ? a : b, a, ntiles, g); });
      std::printf("%-40s mode %d pf %d delay %5d  %d CTAs/SM x %3zu KB smem  %7.3f ms  %7.1f GB/s\n", label, mode, prefetch, delay, o,
                  smem >> 10, ms, (mode == 3 ? 2 : 1) * gb / ms * 1e3);
      std::fflush(stdout);
    };
    if (ctas >= 3) launch(k_tile<3>);
    else if (ctas == 2) launch(k_tile<2>);
    else launch(k_tile<1>);
  };
  if (argc > 2 && argv[2][0] == 'm') {  // the thread mapping of a real pass (pass 5 of the oop plan, low_bits 5)
    const unsigned char in_real[12] = {0, 1, 2, 16, 17, 3, 4, 18, 12, 13, 14, 15};
    const unsigned char out_real[12] = {1, 5, 8, 0, 2, 7, 9, 3, 6, 11, 4, 10};
    const unsigned char in_seq[12] = {0, 1, 2, 3, 4, 16, 17, 18, 12, 13, 14, 15};
    for (int pf : {0, 1}) {
      run("real lanes: ld real, st real", in_real, out_real, false, 3, pf, 2, 73 << 10);
      run("ld real, st sequential", in_real, pats[0].pos, false, 3, pf, 2, 73 << 10);
      run("ld sequential lanes, st real", in_seq, out_real, false, 3, pf, 2, 73 << 10);
      run("ld sequential lanes, st sequential", in_seq, pats[0].pos, false, 3, pf, 2, 73 << 10);
    }
    // where the blocks go: (a) adjacent (above), (b) as the first re-sorting plan placed them (fast
    // tile-number bits -> bits 22-24 and 26-29), (c) fast bits low on both sides (reads lose their adjacency)
    out_runs = {{3, 22}, {4, 26}, {10, 12}, {1, 25}};
    run("ld real, st real, blocks scattered (b)", in_real, out_real, false, 3, 0, 2, 73 << 10);
    run("ld seq, st seq, blocks scattered (b)", in_seq, pats[0].pos, false, 3, 0, 2, 73 << 10);
    run("st only, blocks scattered (b)", in_seq, pats[0].pos, false, 2, 0, 2, 73 << 10);
    out_runs = {{3, 12}, {1, 22}, {1, 15}, {1, 23}, {1, 16}, {1, 24}, {2, 17}, {1, 25}, {1, 19}, {1, 26}, {1, 20}, {1, 27}, {1, 21}, {2, 28}};
    in_runs = {{3, 19}, {1, 5}, {1, 22}, {1, 6}, {1, 23}, {1, 7}, {2, 24}, {1, 8}, {1, 26}, {1, 9}, {1, 27}, {1, 10}, {1, 28}, {1, 11}, {1, 29}};
    run("ld real, st real, fast bits low both sides (c)", in_real, out_real, false, 3, 0, 2, 73 << 10);
    run("ld seq, st seq, fast bits low both sides (c)", in_seq, pats[0].pos, false, 3, 0, 2, 73 << 10);
    run("ld only, (c)", in_seq, pats[0].pos, false, 1, 0, 2, 73 << 10);
    out_runs.clear();
    run("ld seq, st seq, reads as (c), blocks adjacent", in_seq, pats[0].pos, false, 3, 0, 2, 73 << 10);
    in_runs.clear();
    run("ld real only", in_real, out_real, false, 1, 0, 2, 73 << 10);
    run("ld sequential only", in_seq, out_real, false, 1, 0, 2, 73 << 10);
    run("st real only", in_real, out_real, false, 2, 0, 2, 73 << 10);
    run("st sequential only", in_real, pats[0].pos, false, 2, 0, 2, 73 << 10);
    return 0;
  }
  if (argc > 2) {  // focused: the re-sorted layout, prefetch flavours, with and without a per-tile delay
    for (int d : {0, 4000, 8000, 12000}) {
      delay = d;
      for (int pf : {0, 1, 2, 3, 4, 5})
        for (int c : {1, 2}) run("oop: read bits 0-4 + 12-18", pats[9].pos, pats[0].pos, false, 3, pf, c, 73 << 10);
    }
    delay = 8000;
    for (int pf : {0, 1, 3}) run("in place: pass-6-like", pats[5].pos, pats[5].pos, true, 3, pf, 2, 73 << 10);
    for (int pf : {0, 1, 3}) run("in place: 128 B chunks on 64 pages", pats[8].pos, pats[8].pos, true, 3, pf, 2, 73 << 10);
    return 0;
  }
  // 1. shared-memory footprint at 2 CTAs/SM, contiguous tiles and benchmark-like tiles
  for (int p = 0; p < 2; ++p)
    for (size_t kb : {48, 64, 66, 73, 82, 98, 112})
      run(pats[p].name, pats[p].pos, pats[p].pos, true, 3, 1, 2, kb << 10);
  // 2. patterns, in place, 2 CTAs x 73 KB (the specialised kernels today) and x 66 KB
  for (int pi = 0; pi < 4; ++pi)
    for (size_t kb : {73, 66})
      for (int pf : {1, 0}) run(pats[pi].name, pats[pi].pos, pats[pi].pos, true, 3, pf, 2, kb << 10);
  // 3. loads only / stores only per pattern
  for (int pi = 0; pi < 4; ++pi) {
    run(pats[pi].name, pats[pi].pos, pats[pi].pos, true, 1, 1, 2, 73 << 10);
    run(pats[pi].name, pats[pi].pos, pats[pi].pos, true, 2, 0, 2, 73 << 10);
  }
  // 4. out of place: gather reads, contiguous stores
  for (int p = 1; p < 4; ++p)
    for (size_t kb : {73, 66}) {
      char nm[96];
      std::snprintf(nm, sizeof nm, "oop: read %s", pats[p].name);
      run(nm, pats[p].pos, pats[0].pos, false, 3, 1, 2, kb << 10);
    }
  // 4b. how many 2 MB pages does a tile touch?
  for (int p = 4; p < 9; ++p)
    for (int mode : {3, 1, 2}) run(pats[p].name, pats[p].pos, pats[p].pos, true, mode, mode != 2, 2, 73 << 10);
  for (int p = 4; p < 6; ++p) {
    char nm[96];
    std::snprintf(nm, sizeof nm, "oop: read %s", pats[p].name);
    run(nm, pats[p].pos, pats[0].pos, false, 3, 1, 2, 73 << 10);
  }
  // 4c. the re-sorted layouts of the out-of-place passes: gather reads, contiguous block stores
  for (int p = 9; p < 11; ++p) {
    char nm[96];
    std::snprintf(nm, sizeof nm, "oop: read %s", pats[p].name);
    for (int pf : {1, 0}) run(nm, pats[p].pos, pats[0].pos, false, 3, pf, 2, 73 << 10);
    run(nm, pats[p].pos, pats[0].pos, false, 1, 1, 2, 73 << 10);
    run(nm, pats[p].pos, pats[0].pos, false, 3, 1, 1, 73 << 10);
    run(nm, pats[p].pos, pats[0].pos, false, 3, 1, 3, 73 << 10);
  }
  // 5. CTAs per SM at 66 KB each
  for (int c : {1, 2, 3}) run(pats[1].name, pats[1].pos, pats[1].pos, true, 3, 1, c, 66 << 10);
  return 0;
}
