"""Synthetic workloads as primitive op streams (what reaches ``#>`` after qelib1.inc expansion).

Generators for the configurations BASELINE.json names: the QFT in the pattern
of examples/fourier.qasm, random single-qubit/CNOT layers, and the widened ripple-carry adder
of examples/rippleCarryAdder.qasm.  Gate matrices follow the reference's ``unitary`` formula
(QGate.hs:112-118) and its truncated ``pi`` literal (QASM/Simulation.hs:211); every qelib1.inc
gate is expanded to U and CX exactly as the header defines it (examples/qelib1.inc:7-95).
Op stream format: ("U", q, 2x2) | ("CX", c, t), reference qubit indices.
"""
from __future__ import annotations

import numpy as np

from .qgate import unitary_matrix

PI = 3.14159265358979  # QASM/Simulation.hs:211


def u3(q, theta, phi, lam):
    return [("U", q, unitary_matrix(theta, phi, lam))]


def u2(q, phi, lam):
    return u3(q, PI / 2, phi, lam)


def u1(q, lam):
    return u3(q, 0.0, 0.0, lam)


def x(q):
    return u3(q, PI, 0.0, PI)


def h(q):
    return u2(q, 0.0, PI)


def t(q):
    return u1(q, PI / 4)


def tdg(q):
    return u1(q, -PI / 4)


def cx(c, t_):
    return [("CX", c, t_)]


def cu1(lam, a, b):
    """qelib1.inc:78-85."""
    return u1(a, lam / 2) + cx(a, b) + u1(b, -lam / 2) + cx(a, b) + u1(b, lam / 2)


def ccx(a, b, c):
    """qelib1.inc:59-68 (Toffoli = 9 U + 6 CX)."""
    return (h(c) + cx(b, c) + tdg(c) + cx(a, c) + t(c) + cx(b, c) + tdg(c) + cx(a, c) + t(b) + t(c) + h(c)
            + cx(a, b) + t(a) + tdg(b) + cx(a, b))


def qft_ops(n: int, with_x: bool = True):
    """examples/fourier.qasm:8-21 generalised to n qubits: n + 5 n (n-1) / 2 ops (+2 x)."""
    ops = []
    if with_x and n >= 3:
        ops += x(0) + x(2)
    for j in range(n):
        for i in range(j):
            ops += cu1(PI / (2 ** (j - i)), j, i)
        ops += h(j)
    return ops


def random_layers(n: int, depth: int, seed: int = 1000, lam0: bool = True):
    """Layer l (seed + l): U(theta, phi, lambda) on every qubit with angles ~ U[0, 4 pi)
    (test/Qubism/QGateSpec.hs:14-19; lambda = 0 keeps the reference's ``unitary``
    norm-preserving), then CX on the disjoint pairs of a random permutation."""
    ops = []
    for l in range(depth):
        rng = np.random.default_rng(seed + l)
        ang = rng.uniform(0.0, 4.0 * np.pi, size=(n, 3))
        for q in range(n):
            ops.append(("U", q, unitary_matrix(ang[q, 0], ang[q, 1], 0.0 if lam0 else ang[q, 2])))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            ops.append(("CX", int(perm[i]), int(perm[i + 1])))
    return ops


def proper_unitary_layers(n: int, depth: int, seed: int = 3000):
    """Same layout with true SU(2) matrices (OpenQASM's U, all four entries complex and no
    common phase): exercises the GENERAL gate class of the backend."""
    ops = []
    for l in range(depth):
        rng = np.random.default_rng(seed + l)
        ang = rng.uniform(0.0, 4.0 * np.pi, size=(n, 3))
        for q in range(n):
            th, ph, la = ang[q]
            m = np.array([[np.cos(th / 2), -np.exp(1j * la) * np.sin(th / 2)],
                          [np.exp(1j * ph) * np.sin(th / 2), np.exp(1j * (ph + la)) * np.cos(th / 2)]])
            ops.append(("U", q, m))
        perm = rng.permutation(n)
        for i in range(0, n - 1, 2):
            ops.append(("CX", int(perm[i]), int(perm[i + 1])))
    return ops


def adder_ops(k: int):
    """examples/rippleCarryAdder.qasm:6-42 widened to k-bit operands on ONE register
    q[2k+2]: a = q[0..k), b = q[k..2k), cin = q[2k], cout = q[2k+1]."""
    a = lambda i: i
    b = lambda i: k + i
    cin, cout = 2 * k, 2 * k + 1

    def majority(p, q, r):
        return cx(r, q) + cx(r, p) + ccx(p, q, r)

    def unmaj(p, q, r):
        return ccx(p, q, r) + cx(r, p) + cx(p, q)

    ops = x(a(0))
    for i in range(k):
        ops += x(b(i))
    ops += majority(cin, b(0), a(0))
    for i in range(1, k):
        ops += majority(a(i - 1), b(i), a(i))
    ops += cx(a(k - 1), cout)
    for i in reversed(range(1, k)):
        ops += unmaj(a(i - 1), b(i), a(i))
    ops += unmaj(cin, b(0), a(0))
    return ops


def random_mixed(n: int, nops: int, seed: int):
    """A random mix of every op kind the ABI takes -- rotations (lambda = 0), general `unitary`
    matrices, Paulis / Hadamard / a phase gate, CX (dense, so that flip-mask toggles, static and
    masked register swaps all occur) and multi-controlled gates -- for the randomised parity
    sweeps.  Same stream format as the other generators plus ("CU", ctrls, t, 2x2)."""
    rng = np.random.default_rng(seed)
    had = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
    fixed = [had, np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]]),
             np.diag([1, -1]).astype(complex), np.diag([1, 1j])]
    ops = []
    while len(ops) < nops:
        kind = int(rng.integers(0, 8))
        q = int(rng.integers(0, n))
        if kind <= 1:
            ops.append(("U", q, unitary_matrix(float(rng.uniform(0, 12)), float(rng.uniform(0, 12)), 0.0)))
        elif kind == 2:
            ops.append(("U", q, unitary_matrix(*[float(a) for a in rng.uniform(0, 12, 3)])))
        elif kind == 3:
            ops.append(("U", q, fixed[int(rng.integers(0, len(fixed)))]))
        elif kind <= 6:
            c = int(rng.integers(0, n))
            if c != q:
                ops.append(("CX", c, q))
        elif n >= 3:
            cs = [int(a) for a in rng.choice([a for a in range(n) if a != q], int(rng.integers(1, 3)), replace=False)]
            m = unitary_matrix(*[float(a) for a in rng.uniform(0, 12, 3)]) if rng.integers(0, 2) else \
                np.array([[np.cos(.4), -np.sin(.4)], [np.sin(.4), np.cos(.4)]], dtype=complex)
            ops.append(("CU", cs, q, m))
    return ops
