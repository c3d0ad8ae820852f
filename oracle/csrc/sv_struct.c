/* Structured CPU restatement of the qubism state-vector hot path in C + OpenMP.
 * TEST INFRASTRUCTURE and CPU baseline only -- never linked into the product library.
 *
 * One full sweep of the amplitude array per primitive op, no fusion, fp64 complex: the
 * same op-at-a-time schedule the reference evaluator issues (one `#>` per primitive op,
 * src/Qubism/QASM/Simulation.hs:94-122), with the O(4^n)/O(8^n) dense matrices of
 * src/Qubism/QGate.hs:121-154 replaced by their O(2^n) action.
 * Cross-checked against oracle/dense.py (the literal algorithm) in tests/test_oracle.py.
 *
 * PARITY UNPINNED: see oracle/__init__.py.
 *
 * Qubit i of an n-qubit register is bit n-1-i of the amplitude index (StateVec.hs:65-67).
 * Amplitudes are interleaved (re, im) doubles == Storable (Complex Double).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t kind;      /* 0 = U/CU (controlled 1q), 1 = CX, 2 = COLLAPSE, 3 = MEASURE */
  int32_t target;    /* reference qubit index */
  int32_t nctrl;
  int32_t ctrl[4];
  int32_t bit;       /* COLLAPSE: outcome; MEASURE: out, the outcome */
  double m[8];       /* row-major 2x2, (re,im) pairs: a b c d */
  double r;          /* MEASURE: uniform draw; out: pOne */
} sv_op;

int sv_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU baseline sets its thread
 * count explicitly (bench.py --impl reference). */
void sv_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline uint64_t insert_zero(uint64_t x, int pos) {
  uint64_t lo = x & ((1ull << pos) - 1);
  return ((x >> pos) << (pos + 1)) | lo;
}

/* controlled^nctrl (onJust target m) #> v   (QGate.hs:125-132, 148-154, 78-80) */
void sv_apply_1q(double *v, int n, int target, const double *m, const int *ctrl, int nctrl) {
  const int b = n - 1 - target;
  uint64_t cmask = 0;
  for (int i = 0; i < nctrl; i++) cmask |= 1ull << (n - 1 - ctrl[i]);
  const uint64_t half = 1ull << (n - 1), stride = 1ull << b;
  const double ar = m[0], ai = m[1], br = m[2], bi = m[3], cr = m[4], ci = m[5], dr = m[6], di = m[7];
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)half; p++) {
    uint64_t k0 = insert_zero((uint64_t)p, b);
    if ((k0 & cmask) != cmask) continue;
    uint64_t k1 = k0 | stride;
    double x0r = v[2 * k0], x0i = v[2 * k0 + 1], x1r = v[2 * k1], x1i = v[2 * k1 + 1];
    v[2 * k0]     = (ar * x0r - ai * x0i) + (br * x1r - bi * x1i);
    v[2 * k0 + 1] = (ar * x0i + ai * x0r) + (br * x1i + bi * x1r);
    v[2 * k1]     = (cr * x0r - ci * x0i) + (dr * x1r - di * x1i);
    v[2 * k1 + 1] = (cr * x0i + ci * x0r) + (dr * x1i + di * x1r);
  }
}

/* cnot c t #> v  (QGate.hs:121-122): a pure permutation */
void sv_apply_cnot(double *v, int n, int c, int t) {
  const int bt = n - 1 - t;
  const uint64_t cm = 1ull << (n - 1 - c), stride = 1ull << bt, half = 1ull << (n - 1);
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < (int64_t)half; p++) {
    uint64_t k0 = insert_zero((uint64_t)p, bt);
    if (!(k0 & cm)) continue;
    uint64_t k1 = k0 | stride;
    double tr = v[2 * k0], ti = v[2 * k0 + 1];
    v[2 * k0] = v[2 * k1]; v[2 * k0 + 1] = v[2 * k1 + 1];
    v[2 * k1] = tr; v[2 * k1 + 1] = ti;
  }
}

/* S0, S1 = sum |z_k|^2 by the value of qubit q's bit (pOne = sqrt(S1), StateVec.hs:124-126) */
void sv_sumsq(const double *v, int n, int q, double *s0, double *s1) {
  const int b = n - 1 - q;
  const uint64_t N = 1ull << n;
  double a0 = 0.0, a1 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : a0, a1)
  for (int64_t k = 0; k < (int64_t)N; k++) {
    double w = v[2 * k] * v[2 * k] + v[2 * k + 1] * v[2 * k + 1];
    if (((uint64_t)k >> b) & 1) a1 += w; else a0 += w;
  }
  *s0 = a0; *s1 = a1;
}

/* collapse q bit  (StateVec.hs:104-114): mask, then divide by the 2-norm (0 weight -> NaN) */
void sv_collapse(double *v, int n, int q, int bit) {
  const int b = n - 1 - q;
  const uint64_t N = 1ull << n;
  double s0, s1;
  sv_sumsq(v, n, q, &s0, &s1);
  const double nrm = sqrt(bit ? s1 : s0);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)N; k++) {
    int kb = (int)(((uint64_t)k >> b) & 1);
    double keep = (kb == bit) ? 1.0 : 0.0;
    v[2 * k] = (v[2 * k] * keep) / nrm;
    v[2 * k + 1] = (v[2 * k + 1] * keep) / nrm;
  }
}

/* measureQubit q with the draw r supplied  (StateVec.hs:118-129) */
int sv_measure_qubit(double *v, int n, int q, double r, double *pone) {
  double s0, s1;
  sv_sumsq(v, n, q, &s0, &s1);
  double p = s1 > 0.0 ? sqrt(s1) : NAN;
  if (pone) *pone = p;
  int bit = (r < p) ? 1 : 0;
  sv_collapse(v, n, q, bit);
  return bit;
}

void sv_init_basis(double *v, int n) {
  const uint64_t N = 1ull << n;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < (int64_t)(2 * N); k++) v[k] = 0.0;
  v[0] = 1.0;
}

/* run a whole primitive-op stream, one sweep per op */
void sv_run_ops(double *v, int n, sv_op *ops, int64_t nops) {
  for (int64_t i = 0; i < nops; i++) {
    sv_op *o = &ops[i];
    switch (o->kind) {
      case 0: sv_apply_1q(v, n, o->target, o->m, o->ctrl, o->nctrl); break;
      case 1: sv_apply_cnot(v, n, o->ctrl[0], o->target); break;
      case 2: sv_collapse(v, n, o->target, o->bit); break;
      case 3: o->bit = sv_measure_qubit(v, n, o->target, o->r, &o->r); break;
      default: break;
    }
  }
}
