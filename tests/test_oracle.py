"""CPU tests of the oracle itself: the literal-dense restatement (the reference's real
algorithm) against the structured numpy and C restatements, against hand-checkable cases,
against the reference's own QuickCheck properties (test/Qubism/*.hs) and against the
committed golden fixtures.  No GPU, no product code."""
import itertools
import json
import math
import os

import numpy as np
import pytest

from oracle import cport, dense as D, qasm, structured as S

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def c(v):
    a = np.asarray(v, dtype=float)
    return a[:, 0] + 1j * a[:, 1]


def close(a, b, tol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    if np.isnan(a).any() or np.isnan(b).any():  # NaN states are one class
        return bool(np.isnan(a).any() and np.isnan(b).any())
    return bool(np.abs(a - b).max() <= tol)


# ------------------------------------------------------------------ reference quirks
def test_unitary_formula_quirks():
    # QGate.hs:112-118: u1(l) = e^{il/2} I, a pure scalar; h = (i/sqrt2)[[1,1],[-1,1]]
    u1 = D.unitary(0, 0, 0.7)
    assert close(u1, np.exp(0.35j) * np.eye(2))
    pi = 3.14159265358979  # QASM/Simulation.hs:211
    h = D.unitary(pi / 2, 0, pi)
    assert close(h, (1j / math.sqrt(2)) * np.array([[1, 1], [-1, 1]]), 1e-14)
    x = D.unitary(pi, 0, pi)
    assert close(x, np.array([[0, 1j], [-1j, 0]]), 1e-14)
    # not unitary in general: |U^dag U - I| off-diagonal = |sin(lambda) sin(theta)|
    u = D.unitary(0.3, 0.2, 0.1)
    off = abs((u.conj().T @ u)[0, 1])
    assert abs(off - abs(math.sin(0.1) * math.sin(0.3))) < 1e-15


def test_bell_pair_and_dsl_teleportation():
    # examples/Teleportation.hs:21: cnot 0 1 <> onJust 0 hadamard #> |00> = (|00> + |11>)/sqrt2
    bell = D.apply(D.mul(D.cnot(2, 0, 1), D.onJust(2, 0, D.hadamard())), D.mkStateVec(2))
    assert close(bell, np.array([1, 0, 0, 1]) / math.sqrt(2), 1e-15)
    # the whole teleport1 for every outcome pair: qubit 2 ends in Alice's state
    rng = np.random.default_rng(5)
    a = S.gen_state(1, rng)
    for r0, r1 in itertools.product([2.0, -1.0], repeat=2):
        v = D.tensor(a, bell)
        v = D.apply(D.cnot(3, 0, 1), v)
        v = D.apply(D.onJust(3, 0, D.hadamard()), v)
        c0, v, _ = D.measureQubit(3, 0, r0, v)
        c1, v, _ = D.measureQubit(3, 1, r1, v)
        v = D.apply(D.ifBit(3, c0, D.onJust(3, 2, D.pauliZ())), v)
        v = D.apply(D.ifBit(3, c1, D.onJust(3, 2, D.pauliX())), v)
        idx = (c0 << 2) | (c1 << 1)
        assert abs(abs(np.vdot(v[idx:idx + 2], a)) - 1) < 1e-14  # Alice's state, up to a global phase
        assert abs(np.linalg.norm(v) - 1) < 1e-14


def test_pone_is_sqrt_s1_and_zero_weight_collapse_is_nan():
    rng = np.random.default_rng(1)
    v = S.gen_state(3, rng)
    for q in range(3):
        _, _, p = D.measureQubit(3, q, 0.5, v)
        assert abs(p - math.sqrt(S.sumsq(3, q, v)[1])) < 1e-15
    z = D.mkStateVec(2)  # |00>: qubit 0 is never One
    with np.errstate(all="ignore"):
        assert np.isnan(D.collapse(2, 0, 1, z)).all()
        assert np.isnan(S.collapse(2, 0, 1, z)).all()
    bit, v2, p = D.measureQubit(2, 0, 0.0, z)  # pOne = NaN -> r < NaN False -> Zero
    assert bit == 0 and math.isnan(p) and close(v2, z)


def test_controlled_is_literal_formula():
    # QGate.hs:125-132 on a gate that touches the control qubit itself: M.P + I - P literally
    m = D.onJust(2, 0, D.pauliX())
    P = np.diag([0, 0, 1, 1]).astype(complex)
    assert close(D.controlled(2, 0, m), m @ P + np.eye(4) - P)


# ------------------------------------------------------------------ restatements agree
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
def test_dense_structured_cport_agree(n):
    rng = np.random.default_rng(100 + n)
    v = S.gen_state(n, rng)
    for q in range(n):
        m = D.unitary(*rng.uniform(0, 4 * np.pi, 3))
        ref = D.apply(D.onJust(n, q, m), v)
        assert close(ref, S.apply_1q(n, q, m, v), 1e-14)
        assert close(ref, cport.run_ops(n, [("U", q, m)], v), 1e-14)
        for b in (0, 1):
            ref = D.collapse(n, q, b, v)
            assert close(ref, S.collapse(n, q, b, v), 1e-14)
            assert close(ref, cport.run_ops(n, [("COLLAPSE", q, b)], v), 1e-14)
        for ct in range(n):
            if ct == q:
                continue
            ref = D.apply(D.cnot(n, ct, q), v)
            assert close(ref, S.apply_cnot(n, ct, q, v), 0)
            assert close(ref, cport.run_ops(n, [("CX", ct, q)], v), 0)
            ref = D.apply(D.controlled(n, ct, D.onJust(n, q, m)), v)
            assert close(ref, S.apply_1q(n, q, m, v, ctrls=(ct,)), 1e-14)
            assert close(ref, cport.run_ops(n, [("CU", [ct], q, m)], v), 1e-14)
    if n >= 3:
        ref = D.apply(D.controlled(n, 0, D.controlled(n, 2, D.onJust(n, 1, m))), v)
        assert close(ref, S.apply_1q(n, 1, m, v, ctrls=(0, 2)), 1e-14)
        M = D.kronecker(D.unitary(1, 2, 3), D.unitary(.4, .5, .6))
        full = D.kronecker(D.kronecker(np.eye(1 << 1), M), np.eye(1 << (n - 3))) if n > 3 else D.kronecker(np.eye(2), M)
        assert close(D.apply(full, v), S.apply_kq(n, [1, 2], M, v), 1e-14)
        assert close(D.apply(D.onRange(n, 0, 2, m), v),
                     S.run_ops(n, [("U", 2, m), ("U", 1, m), ("U", 0, m)], v), 1e-13)
        assert close(D.apply(D.onEvery(n, m), v), S.run_ops(n, [("U", i, m) for i in range(n)], v), 1e-13)


def test_measure_paths_agree():
    rng = np.random.default_rng(3)
    n = 4
    v = S.gen_state(n, rng)
    for rs in itertools.product([2.0, -1.0, 0.5], repeat=n):
        bits, out = D.measure(n, list(rs), v)
        w = v
        got = []
        for q in range(n):
            b, w, _ = S.measure_qubit(n, q, rs[q], w)
            got.append(b)
        assert bits == got and close(out, w, 1e-13)
        ops = [("MEASURE", q, rs[q]) for q in range(n)]
        assert close(out, cport.run_ops(n, ops, v), 1e-13)


# ------------------------------------------------------------------ the reference's properties
def test_quickcheck_vector_space_and_hilbert_laws():
    # test/Qubism/AlgebraTests.hs:25-47 on StateVec 1 (StateVecSpec.hs:49-50) and wider
    rng = np.random.default_rng(11)
    for n in (1, 3):
        for _ in range(50):
            a, b, w = (S.gen_state(n, rng) for _ in range(3))
            z = complex(*rng.uniform(-1, 1, 2))
            assert D.approx_eq((a + b) + w, a + (b + w)) and D.approx_eq(a + b, b + a)
            assert D.approx_eq(-a + a, D.zero(n)) and D.approx_eq(z * (a + b), z * a + z * b)
            assert abs(D.inner(w, z * a + b) - (z * D.inner(w, a) + D.inner(w, b))) < 1e-5
            assert D.inner(a, b) == D.inner(b, a).conjugate()  # exact (AlgebraTests.hs:43-47)


def test_quickcheck_measurement_idempotent():
    # StateVecSpec.hs:35-62: with the same draws, st >> st == st
    rng = np.random.default_rng(12)
    for _ in range(100):
        v = S.gen_state(1, rng)
        r = rng.uniform()
        b1, v1, _ = D.measureQubit(1, 0, r, v)
        b2, v2, _ = D.measureQubit(1, 0, r, v1)
        assert b1 == b2 and D.approx_eq(v1, v2)


# ------------------------------------------------------------------ interpreter + golden fixtures
def test_expr_and_parser_quirks():
    ast = qasm.parse('qreg q[1]; U(-pi/2, 2 pow 3, sin 1 + 2*3) q[0];')
    u = ast[1][2][1]
    assert qasm.eval_expr(u[1]) == -3.14159265358979 / 2
    assert qasm.eval_expr(u[2]) == 8.0
    assert abs(qasm.eval_expr(u[3]) - (math.sin(1) + 6)) < 1e-15
    with pytest.raises(qasm.ParseError):
        qasm.parse("qreg q[1]; h q[0];")  # h undeclared without the include
    with pytest.raises(qasm.RuntimeErrorQ):
        qasm.run_qasm('include "qelib1.inc"; qreg a[2]; qreg b[3]; cx a,b;')


@pytest.mark.parametrize("name", ["teleportation", "fourier4", "invqft4", "adder2"])
def test_golden_programs(name):
    entry = GOLD[name]
    assert any(not r["degenerate"] for r in entry["runs"])
    for run in entry["runs"]:
        if run["degenerate"]:  # forced onto a zero-weight branch: rounding noise, not comparable
            continue
        for backend in (qasm.DenseBackend(), qasm.StructuredBackend()):
            ps = qasm.run_qasm(entry["source"], backend=backend, draws=run["draws"])
            assert ps.cregs == run["cregs"]
            assert set(ps.stVecs) == set(run["states"])
            for k, v in run["states"].items():
                assert close(ps.stVecs[k][1], c(v), 1e-12), (name, run["draws"], k)


def test_withindex_writeback_bug_is_reproduced():
    # Simulation.hs:101: after two qregs fuse, 1-qubit gates on name[k] land in an orphan map
    # entry; the ripple-carry adder therefore computes a different state than intended.
    buggy = GOLD["adder2"]["runs"][-1]
    fixed = GOLD["adder2_fixed"]["runs"][0]
    # (neither adds correctly: under the reference's `unitary` formula t/tdg are pure scalars,
    #  so its ccx is not a Toffoli -- SURVEY.md section 0 item 5)
    assert buggy["cregs"]["ans"] != fixed["cregs"]["ans"]
    assert len(buggy["states"]) > len(fixed["states"])  # orphans left behind
    live = [k for k in buggy["states"] if "(x)" in k and k.count("(x)") == 3][0]
    assert not close(c(buggy["states"][live]), c(fixed["states"][live]), 1e-6)
    # the op stream that reaches the live state has no U ops after the first fusion
    us = [t for t in buggy["trace"] if t[0] == "U" and "(x)" in str(t[1])]
    assert us == []


def test_golden_gate_vectors():
    for case in GOLD["gate_vectors"]:
        n, op, vin, vout = case["n"], case["op"], c(case["in"]), c(case["out"])
        if op[0] in ("U", "CU"):
            m = D.unitary(*op[-1]["angles"])
            sop = ("U", op[1], m) if op[0] == "U" else ("CU", op[1], op[2], m)
        else:
            sop = tuple(op)
        with np.errstate(all="ignore"):
            assert close(S.run_ops(n, [sop], vin), vout, 1e-13)
            assert close(cport.run_ops(n, [sop], vin), vout, 1e-13)


# ------------------------------------------------------------------ the pin: the reference itself
REF_PATH = os.path.join(os.path.dirname(__file__), "golden", "reference.json")


def test_oracle_against_the_reference_when_it_has_been_run():
    """tests/golden/reference.json holds what the UNMODIFIED reference computes on the golden cases
    (oracle/ref_haskell/run_reference.sh on a box with stack / GHC 8.4).  The build image has no
    Haskell toolchain, so the file may be missing: parity is then 'unpinned' 
    and this test says so instead of passing silently."""
    if not os.path.exists(REF_PATH):
        pytest.skip("parity unpinned: the reference has not been run (oracle/ref_haskell/README.md)")
    ref = json.load(open(REF_PATH))
    assert len(ref["gate_vectors"]) == len(GOLD["gate_vectors"])
    for i, case in enumerate(GOLD["gate_vectors"]):
        want = c(ref["gate_vectors"][str(i)])
        with np.errstate(all="ignore"):
            assert close(c(case["out"]), want, 1e-12), ("gate vector", i, case["op"])
    for name in ("teleportation", "fourier4", "invqft4", "adder2"):
        for r, run in enumerate(GOLD[name]["runs"]):
            got = ref["programs"][name][str(r)]
            assert got["error"] is None, got["error"]
            if run["degenerate"]:
                continue
            assert got["cregs"] == run["cregs"], (name, r)
            assert set(got["states"]) == set(run["states"])
            for k, v in run["states"].items():
                assert close(c(got["states"][k]), c(v), 1e-12), (name, r, k)
