/* A host that is neither Python nor C++: plain C99 over include/qubism_sv.h, making the calls the
 * Haskell shim's `foreign import ccall` stubs make.
 *
 *   1. examples/Teleportation.hs:14-31 -- teleport1: Alice's qubit `tensor` a Bell pair, cnot 0 1,
 *      hadamard on 0, measureQubit 0 and 1 with the draws passed in, ifBit corrections on qubit 2;
 *      every gate goes through the interpreter's value-semantic pattern (QASM/Simulation.hs:94-122):
 *      sv' = g #> sv  ==  qb_state_apply_pure(sv, op, &sv'), then the old value is dropped
 *      (qb_state_free).  Check: qubit 2 carries Alice's amplitudes for all four outcome pairs.
 *   2. QGate.hs:112-118 -- `unitary theta phi lambda` evaluated here in C exactly as the reference
 *      writes it, on every qubit of a 12-qubit register, against the closed form of the product
 *      state (each amplitude is a product of matrix entries of column 0).
 *
 * Exit code 0 = everything matched; 3 = no CUDA device (qb_init fails loudly: there is no CPU path);
 * anything else = a mismatch or an error (text on stderr).  Built and run by tests/test_chost.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "qubism_sv.h"

#define CHECK(call)                                                                       \
  do {                                                                                    \
    int rc_ = (call);                                                                     \
    if (rc_ != QB_OK) {                                                                   \
      fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, qb_last_error()); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

static qb_c64 c(double re, double im) {
  qb_c64 z;
  z.re = re;
  z.im = im;
  return z;
}
static qb_c64 cmul(qb_c64 a, qb_c64 b) { return c(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static qb_c64 cexpi(double t) { return c(cos(t), sin(t)); }

/* QGate.hs:112-118, `unitary theta phi lambda` as the reference writes it (row-major a b c d;
 * a = d = cis (phi + lambda/2) cos (theta/2): not the OpenQASM U and not unitary in general) */
static void unitary(double th, double ph, double la, qb_c64 m[4]) {
  const double cs = cos(th / 2), sn = sin(th / 2);
  m[0] = cmul(cexpi(ph + la / 2), c(cs, 0));
  m[1] = cmul(cexpi(ph - la / 2), c(-sn, 0));
  m[2] = cmul(cexpi(ph - la / 2), c(sn, 0));
  m[3] = cmul(cexpi(ph + la / 2), c(cs, 0));
}

static qb_op gate_op(int target, const qb_c64 m[4]) {
  qb_op o;
  int i;
  o.kind = 0;
  o.target = target;
  o.nctrl = 0;
  for (i = 0; i < 4; ++i) o.ctrl[i] = 0;
  o._pad = 0;
  for (i = 0; i < 4; ++i) o.m[i] = m[i];
  return o;
}
static qb_op cnot_op(int ctrl, int target) {
  const qb_c64 x[4] = {{0, 0}, {1, 0}, {1, 0}, {0, 0}};
  qb_op o = gate_op(target, x);
  o.kind = 1;
  o.nctrl = 1;
  o.ctrl[0] = ctrl;
  return o;
}

/* sv <- g #> sv, the interpreter's way: a new value, the old one dropped */
static int pure_step(qb_state **sv, qb_op op) {
  qb_state *next = NULL;
  CHECK(qb_state_apply_pure(*sv, &op, 1, &next));
  qb_state_free(*sv);
  *sv = next;
  return 0;
}

static int teleport(qb_ctx *ctx, double r0, double r1, double *worst) {
  const qb_c64 H[4] = {{0.70710678118654752, 0}, {0.70710678118654752, 0}, {0.70710678118654752, 0}, {-0.70710678118654752, 0}};
  const qb_c64 Z[4] = {{1, 0}, {0, 0}, {0, 0}, {-1, 0}};
  const qb_c64 X[4] = {{0, 0}, {1, 0}, {1, 0}, {0, 0}};
  qb_c64 alice[2], out[8];
  qb_state *a = NULL, *pair = NULL, *total = NULL;
  int c0 = -1, c1 = -1, i;
  double p;
  alice[0] = c(0.6, 0.0);
  alice[1] = c(0.0, 0.8); /* 0.6 |0> + 0.8i |1> */
  CHECK(qb_state_from_host(ctx, 1, alice, &a));
  CHECK(qb_state_create(ctx, 2, 1, &pair));
  if (pure_step(&pair, gate_op(0, H)) || pure_step(&pair, cnot_op(0, 1))) return 1;
  CHECK(qb_tensor(a, pair, &total));
  qb_state_free(a);
  qb_state_free(pair);
  if (pure_step(&total, cnot_op(0, 1)) || pure_step(&total, gate_op(0, H))) return 1;
  CHECK(qb_measure_qubit(total, 0, r0, &c0, &p));
  CHECK(qb_measure_qubit(total, 1, r1, &c1, &p));
  if (c0 && pure_step(&total, gate_op(2, Z))) return 1;
  if (c1 && pure_step(&total, gate_op(2, X))) return 1;
  CHECK(qb_state_read(total, 0, 8, out));
  qb_state_free(total);
  for (i = 0; i < 8; ++i) { /* index = q0 q1 q2 (qubit 0 is the most significant bit, StateVec.hs:65-67) */
    const int here = ((i >> 2) & 1) == c0 && ((i >> 1) & 1) == c1;
    /* (the reference applies Z before X: for c0 = c1 = 1 that is X Z = -(Z X), a global phase of -1) */
    const double sign = (c0 && c1) ? -1.0 : 1.0;
    const qb_c64 want = here ? c(sign * alice[i & 1].re, sign * alice[i & 1].im) : c(0, 0);
    const double d = hypot(out[i].re - want.re, out[i].im - want.im);
    if (d > *worst) *worst = d;
  }
  return 0;
}

static int product_state(qb_ctx *ctx, double *worst) {
  enum { N = 12 };
  qb_c64 m[N][4], *out;
  qb_state *sv = NULL;
  int q;
  uint64_t i;
  CHECK(qb_state_create(ctx, N, 1, &sv));
  for (q = 0; q < N; ++q) {
    unitary(0.3 + 0.37 * q, 1.1 - 0.21 * q, 0.05 * q, m[q]);
    if (pure_step(&sv, gate_op(q, m[q]))) return 1;
  }
  out = (qb_c64 *)malloc(sizeof(qb_c64) << N);
  if (!out) return 1;
  CHECK(qb_state_read(sv, 0, (uint64_t)1 << N, out));
  qb_state_free(sv);
  for (i = 0; i < ((uint64_t)1 << N); ++i) {
    qb_c64 want = c(1, 0);
    double d;
    for (q = 0; q < N; ++q) want = cmul(want, m[q][((i >> (N - 1 - q)) & 1) ? 2 : 0]); /* column 0 of each gate */
    d = hypot(out[i].re - want.re, out[i].im - want.im);
    if (d > *worst) *worst = d;
  }
  free(out);
  return 0;
}

int main(void) {
  qb_ctx *ctx = NULL;
  double worst = 0.0;
  const double draws[4][2] = {{0.1, 0.1}, {0.1, 0.9}, {0.9, 0.1}, {0.9, 0.9}};
  qb_stats st;
  int k, rc = qb_init(0, &ctx);
  if (rc != QB_OK) {
    fprintf(stderr, "qb_init -> %d: %s\n", rc, qb_last_error());
    return 3;
  }
  for (k = 0; k < 4; ++k)
    if (teleport(ctx, draws[k][0], draws[k][1], &worst)) return 1;
  if (product_state(ctx, &worst)) return 1;
  CHECK(qb_get_stats(ctx, &st));
  printf("%s: max |difference| %.3g, %llu ops submitted, %llu lazy clones, %llu separate copies\n", qb_version(), worst,
         (unsigned long long)st.ops_submitted, (unsigned long long)st.clones, (unsigned long long)st.cow_copies);
  CHECK(qb_shutdown(ctx));
  if (!(worst < 1e-12)) {
    fprintf(stderr, "mismatch: %.3g\n", worst);
    return 2;
  }
  printf("ok\n");
  return 0;
}
