// Prelude of the structure-specialised fused pass (qb_jit.cpp).
//
// The step kernels of qb_kernels.cu read the STRUCTURE of a pass (which register bit carries which
// tile bit in which round, which slot of which step holds which gate class, toggles, register
// swaps) from the constant bank at run time and are therefore bound by instruction issue: skip
// branches, header loads, XOR-swap register moves, and -- because every gate arm must leave every
// amplitude in the register it came in -- three DFMAs per rotation where two would do.  For a pass
// structure that comes back (an iterated circuit), qb_jit.cpp emits the same pass as straight-line
// code: structure as literals, gate COEFFICIENTS still run-time kernel parameters.  The compiler
// then renames registers instead of moving data:
//   * a rotation is c * [[1,-t],[t,1]] (|t| <= 1; the cosine goes to the state's deferred scalar):
//     2 DFMA per component pair, the new x0 simply lives in a fresh register;
//   * an X / CX whose control is a register bit no flip can have touched is a renaming: 0 instructions;
//   * slot kinds, toggle masks, shared-memory offsets and global strides are immediates.
// This file is compiled by NVRTC for sm_100a (device) and, with QB_JIT_HOST, by g++ for the
// CPU-side emulation the tests use to check the generator without a GPU.
#ifdef QB_JIT_HOST
#include <cmath>
#include <cstdint>
#include <cstring>
#define QBJ_DEV static inline
typedef uint32_t u32;
typedef uint64_t u64;
typedef uint16_t u16;
static inline double qbj_fma(double a, double b, double c) { return std::fma(a, b, c); }
static inline double qbj_sign_xor(double v, u32 bit) {
  u64 x;
  std::memcpy(&x, &v, 8);
  x ^= (u64)(bit & 1u) << 63;
  std::memcpy(&v, &x, 8);
  return v;
}
static inline void qbj_cswap(double &a, double &b, bool ok) {
  if (ok) {
    const double t = a;
    a = b;
    b = t;
  }
}
#else
#define QBJ_DEV __device__ __forceinline__
typedef unsigned int u32;
typedef unsigned long long u64;
typedef unsigned short u16;
QBJ_DEV double qbj_fma(double a, double b, double c) { return fma(a, b, c); }
QBJ_DEV double qbj_sign_xor(double v, u32 bit) {
  return __longlong_as_double(__double_as_longlong(v) ^ (long long)((u64)(bit & 1u) << 63));
}
// conditional swap with the masked-XOR trick: no branch, no select (a predicated arm would force
// the renamed registers back into place with moves)
QBJ_DEV void qbj_cswap(double &a, double &b, bool ok) {
  const long long m = ok ? -1ll : 0ll;
  const long long x = (__double_as_longlong(a) ^ __double_as_longlong(b)) & m;
  a = __longlong_as_double(__double_as_longlong(a) ^ x);
  b = __longlong_as_double(__double_as_longlong(b) ^ x);
}
// global accesses: cache policy selected by QBJ_MEM (option jit_mem; measured in profiles/)
//   0 ld.cs / st.cs   1 ld.cs / st.wb   2 ld.ca-default / st.cs   3 ld.cs / st.cg   4 ld.lu / st.cs
//   5 L2 evict_first on both   6 ld.cs / st with L2 evict_last   7 ld.cs / st with L2 evict_first
#ifndef QBJ_MEM
#define QBJ_MEM 0
#endif
#if QBJ_MEM == 1
#define QBJ_LDQ "ld.global.cs"
#define QBJ_STQ "st.global.wb"
#elif QBJ_MEM == 2
#define QBJ_LDQ "ld.global"
#define QBJ_STQ "st.global.cs"
#elif QBJ_MEM == 3
#define QBJ_LDQ "ld.global.cs"
#define QBJ_STQ "st.global.cg"
#elif QBJ_MEM == 4
#define QBJ_LDQ "ld.global.lu"
#define QBJ_STQ "st.global.cs"
#else
#define QBJ_LDQ "ld.global.cs"
#define QBJ_STQ "st.global.cs"
#endif
#if QBJ_MEM >= 5
#define QBJ_HINT 1
QBJ_DEV u64 qbj_policy_first() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
QBJ_DEV u64 qbj_policy_last() {
  u64 p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
#endif
QBJ_DEV void qbj_ld128(const double2 *p, double &a0, double &a1) {
#if QBJ_MEM == 5
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(a0), "=d"(a1) : "l"(p), "l"(qbj_policy_first()));
#else
  asm volatile(QBJ_LDQ ".v2.f64 {%0,%1}, [%2];" : "=d"(a0), "=d"(a1) : "l"(p));
#endif
}
QBJ_DEV void qbj_st128(double2 *p, double a0, double a1) {
#if QBJ_MEM == 5 || QBJ_MEM == 7
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a0), "d"(a1), "l"(qbj_policy_first()) : "memory");
#elif QBJ_MEM == 6
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a0), "d"(a1), "l"(qbj_policy_last()) : "memory");
#else
  asm volatile(QBJ_STQ ".v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a0), "d"(a1) : "memory");
#endif
}
#ifndef QBJ_NO_LD256
QBJ_DEV void qbj_ld256(const double2 *p, double &a0, double &a1, double &b0, double &b1) {
#if QBJ_MEM == 5
  asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
               : "=d"(a0), "=d"(a1), "=d"(b0), "=d"(b1)
               : "l"(p), "l"(qbj_policy_first()));
#else
  asm volatile(QBJ_LDQ ".v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(b0), "=d"(b1) : "l"(p));
#endif
}
QBJ_DEV void qbj_st256(double2 *p, double a0, double a1, double b0, double b1) {
#if QBJ_MEM == 5 || QBJ_MEM == 7
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a0), "d"(a1), "d"(b0), "d"(b1),
               "l"(qbj_policy_first())
               : "memory");
#elif QBJ_MEM == 6
  asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a0), "d"(a1), "d"(b0), "d"(b1),
               "l"(qbj_policy_last())
               : "memory");
#else
  asm volatile(QBJ_STQ ".v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a0), "d"(a1), "d"(b0), "d"(b1) : "memory");
#endif
}
#else  // an NVRTC older than CUDA 12.9: two 128-bit accesses
QBJ_DEV void qbj_ld256(const double2 *p, double &a0, double &a1, double &b0, double &b1) {
  qbj_ld128(p, a0, a1);
  qbj_ld128(p + 1, b0, b1);
}
QBJ_DEV void qbj_st256(double2 *p, double a0, double a1, double b0, double b1) {
  qbj_st128(p, a0, a1);
  qbj_st128(p + 1, b0, b1);
}
#endif
#endif

#ifndef QB_JIT_HOST
// ---- bulk-asynchronous copies (the TMA engine without a tensor map: cp.async.bulk) + mbarrier ----
QBJ_DEV void qbj_mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
QBJ_DEV void qbj_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
QBJ_DEV void qbj_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one arrival that also announces `bytes` of copy traffic: the phase completes when they have landed
QBJ_DEV void qbj_mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
QBJ_DEV void qbj_mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "QBJ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra QBJ_DONE;\n"
      "bra QBJ_WAIT;\n"
      "QBJ_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared, `bytes` (multiple of 16) contiguous, completion counted on the mbarrier
QBJ_DEV void qbj_bulk_load(u32 dst, const void *src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
#endif

#define QBJ_NR (1 << QBJ_R)
#define QBJ_PAIRS(J)                                                              \
  _Pragma("unroll") for (int p = 0; p < (QBJ_NR >> 1); ++p)                       \
    if (const int i0 = ((p >> (J)) << ((J) + 1)) | (p & ((1 << (J)) - 1)); true)  \
      if (const int i1 = i0 | (1 << (J)); true)

QBJ_DEV u32 qbj_swz(u32 u) { return u ^ ((u >> 3) & 7u) ^ ((u >> 6) & 7u) ^ ((u >> 9) & 7u) ^ ((u >> 12) & 7u); }

// ---- rotations [[c,-s],[s,c]], c >= 0, c^2 + s^2 = 1 --------------------------------------------
// form A (|s| <= c): c * [[1,-t],[t,1]], t = s / c.   The factor c is NOT applied here.
template <int J>
QBJ_DEV void qbj_rot_a(double (&re)[QBJ_NR], double (&im)[QBJ_NR], double t) {
  QBJ_PAIRS(J) {
    const double nr = qbj_fma(-t, re[i1], re[i0]);
    const double ni = qbj_fma(-t, im[i1], im[i0]);
    re[i1] = qbj_fma(t, re[i0], re[i1]);
    im[i1] = qbj_fma(t, im[i0], im[i1]);
    re[i0] = nr;
    im[i0] = ni;
  }
}
// the same where a flip may be pending on bit J: the pair is held in reverse order, which turns
// the rotation into the one by the opposite angle (t -> -t; the cosine is unchanged)
template <int J>
QBJ_DEV void qbj_rot_a_flip(double (&re)[QBJ_NR], double (&im)[QBJ_NR], double t, u32 f) {
  qbj_rot_a<J>(re, im, qbj_sign_xor(t, f >> J));
}
// form B (|s| > c): s * [[u,-1],[1,u]], u = c / s.   The factor s is NOT applied here.
template <int J>
QBJ_DEV void qbj_rot_b(double (&re)[QBJ_NR], double (&im)[QBJ_NR], double u) {
  QBJ_PAIRS(J) {
    const double nr = qbj_fma(u, re[i0], -re[i1]);
    const double ni = qbj_fma(u, im[i0], -im[i1]);
    re[i1] = qbj_fma(u, re[i1], re[i0]);
    im[i1] = qbj_fma(u, im[i1], im[i0]);
    re[i0] = nr;
    im[i0] = ni;
  }
}
// three shears (exact rotation, no deferred factor): the flip-aware flavour for |s| > c, where
// the reversed pair order would change the SIGN of form B's deferred factor per thread
template <int J>
QBJ_DEV void qbj_rot3_flip(double (&re)[QBJ_NR], double (&im)[QBJ_NR], double t, double s, u32 f) {
  const double tv = qbj_sign_xor(t, f >> J), sv = qbj_sign_xor(s, f >> J);
  QBJ_PAIRS(J) {
    re[i0] = qbj_fma(tv, re[i1], re[i0]);
    im[i0] = qbj_fma(tv, im[i1], im[i0]);
  }
  QBJ_PAIRS(J) {
    re[i1] = qbj_fma(sv, re[i0], re[i1]);
    im[i1] = qbj_fma(sv, im[i0], im[i1]);
  }
  QBJ_PAIRS(J) {
    re[i0] = qbj_fma(tv, re[i1], re[i0]);
    im[i0] = qbj_fma(tv, im[i1], im[i0]);
  }
}

// ---- real 2x2 [[a,b],[c,d]] (m = a b c d) --------------------------------------------------------
template <int J, int FL>
QBJ_DEV void qbj_real(double (&re)[QBJ_NR], double (&im)[QBJ_NR], const double *m, u32 f) {
  double a = m[0], b = m[1], c = m[2], d = m[3];
  if (FL != 0 && ((f >> J) & 1u)) {
    double t;
    t = a; a = d; d = t;
    t = b; b = c; c = t;
  }
  QBJ_PAIRS(J) {
    const double nr = qbj_fma(a, re[i0], b * re[i1]);
    const double ni = qbj_fma(a, im[i0], b * im[i1]);
    re[i1] = qbj_fma(d, re[i1], c * re[i0]);
    im[i1] = qbj_fma(d, im[i1], c * im[i0]);
    re[i0] = nr;
    im[i0] = ni;
  }
}

// ---- complex 2x2, row-major (re,im): a b c d ------------------------------------------------------
template <int J, int FL>
QBJ_DEV void qbj_general(double (&re)[QBJ_NR], double (&im)[QBJ_NR], const double *m, u32 f) {
  double Ar = m[0], Ai = m[1], Br = m[2], Bi = m[3];
  double Cr = m[4], Ci = m[5], Dr = m[6], Di = m[7];
  if (FL != 0 && ((f >> J) & 1u)) {
    double t;
    t = Ar; Ar = Dr; Dr = t;
    t = Ai; Ai = Di; Di = t;
    t = Br; Br = Cr; Cr = t;
    t = Bi; Bi = Ci; Ci = t;
  }
  QBJ_PAIRS(J) {
    double P = -Ai * im[i0];
    double Q = Ai * re[i0];
    double Tr = Cr * re[i0];
    double Ti = Cr * im[i0];
    P = qbj_fma(Br, re[i1], P);
    Q = qbj_fma(Br, im[i1], Q);
    Tr = qbj_fma(-Ci, im[i0], Tr);
    Ti = qbj_fma(Ci, re[i0], Ti);
    P = qbj_fma(-Bi, im[i1], P);
    Q = qbj_fma(Bi, re[i1], Q);
    Tr = qbj_fma(-Di, im[i1], Tr);
    Ti = qbj_fma(Di, re[i1], Ti);
    re[i0] = qbj_fma(Ar, re[i0], P);
    im[i0] = qbj_fma(Ar, im[i0], Q);
    re[i1] = qbj_fma(Dr, re[i1], Tr);
    im[i1] = qbj_fma(Dr, im[i1], Ti);
  }
}
// complex 2x2 with m00 = 1 (no flip pending): 12 operations per pair
template <int J>
QBJ_DEV void qbj_general1(double (&re)[QBJ_NR], double (&im)[QBJ_NR], const double *m) {
  const double Br = m[2], Bi = m[3], Cr = m[4], Ci = m[5], Dr = m[6], Di = m[7];
  QBJ_PAIRS(J) {
    double Tr = Cr * re[i0];
    double Ti = Cr * im[i0];
    const double P = qbj_fma(Br, re[i1], re[i0]);
    const double Q = qbj_fma(Br, im[i1], im[i0]);
    Tr = qbj_fma(-Ci, im[i0], Tr);
    Ti = qbj_fma(Ci, re[i0], Ti);
    Tr = qbj_fma(-Di, im[i1], Tr);
    Ti = qbj_fma(Di, re[i1], Ti);
    re[i0] = qbj_fma(-Bi, im[i1], P);
    im[i0] = qbj_fma(Bi, re[i1], Q);
    re[i1] = qbj_fma(Dr, re[i1], Tr);
    im[i1] = qbj_fma(Dr, im[i1], Ti);
  }
}

// ---- X / CX -----------------------------------------------------------------------------------------
// every pair along bit J swaps in every thread (uncontrolled X): a renaming
template <int J>
QBJ_DEV void qbj_swap_all(double (&re)[QBJ_NR], double (&im)[QBJ_NR]) {
  QBJ_PAIRS(J) {
    double t = re[i0]; re[i0] = re[i1]; re[i1] = t;
    t = im[i0]; im[i0] = im[i1]; im[i1] = t;
  }
}
// the pairs whose register index has bit K set swap, in every thread (CX, control = register bit K
// that no flip can have touched): a renaming
template <int J, int K>
QBJ_DEV void qbj_swap_static(double (&re)[QBJ_NR], double (&im)[QBJ_NR]) {
  QBJ_PAIRS(J) {
    if ((i0 >> K) & 1) {
      double t = re[i0]; re[i0] = re[i1]; re[i1] = t;
      t = im[i0]; im[i0] = im[i1]; im[i1] = t;
    }
  }
}
// general flavour: register-bit controls under a possibly pending flip, thread / external controls
template <int J>
QBJ_DEV void qbj_swap_dyn(double (&re)[QBJ_NR], double (&im)[QBJ_NR], u32 creg, bool okt, u32 f) {
  QBJ_PAIRS(J) {
    const bool ok = okt && ((((u32)i0 ^ f) & creg) == creg);
    qbj_cswap(re[i0], re[i1], ok);
    qbj_cswap(im[i0], im[i1], ok);
  }
}

// deferred complex scalar on the way out
QBJ_DEV void qbj_scale(double (&re)[QBJ_NR], double (&im)[QBJ_NR], double sr, double si) {
  if (si == 0.0) {
    if (sr != 1.0) {  // (exactly 1: nothing pending)
#pragma unroll
      for (int i = 0; i < QBJ_NR; ++i) {
        re[i] *= sr;
        im[i] *= sr;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < QBJ_NR; ++i) {
      const double xr = re[i], xi = im[i];
      re[i] = qbj_fma(sr, xr, -si * xi);
      im[i] = qbj_fma(sr, xi, si * xr);
    }
  }
}
