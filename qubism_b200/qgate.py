"""Host-side mirror of ``Qubism.QGate`` (src/Qubism/QGate.hs:14-31) over the C ABI.

The reference's ``QGate n`` is a dense 2^n x 2^n matrix value.  That cannot exist at 30
qubits, so here a ``QGate`` is SYMBOLIC: a sum of terms, each a coefficient times a product
of primitive factors (controlled 1-qubit gates and small dense blocks).  The algebra of the
reference module is kept -- ``<>`` (``@``), ``mempty`` (``ident``), ``.:`` / ``+:`` / ``neg``
(VectorSpace, QGate.hs:64-68), ``*:`` (Algebra, :70-71), ``kronecker``, ``controlled``,
``ifBit``, ``onJust``, ``onEvery``, ``onRange`` -- and ``apply`` (``#>``) streams the factors
through the C ABI, where they are fused into few passes over the state.
"""
from __future__ import annotations

import math

import numpy as np

from .statevec import StateVec

C128 = np.complex128


class _Factor:
    __slots__ = ("qs", "m", "ctrls")

    def __init__(self, qs, m, ctrls=()):
        self.qs = tuple(qs)          # target qubits, qs[0] most significant index bit of m
        self.m = np.asarray(m, dtype=C128).reshape(1 << len(self.qs), 1 << len(self.qs))
        self.ctrls = tuple(ctrls)

    def shifted(self, d):
        return _Factor([q + d for q in self.qs], self.m, [c + d for c in self.ctrls])


class QGate:
    """``QGate n``: sum_k coef_k * (product of factors), factors listed in APPLICATION order."""

    def __init__(self, n: int, terms=None):
        self.n = n
        self.terms = terms if terms is not None else [(1.0 + 0j, [])]

    # Semigroup / Monoid (QGate.hs:58-62): (a <> b) #> v = a #> (b #> v)
    def __matmul__(self, other: "QGate") -> "QGate":
        _same(self, other)
        return QGate(self.n, [(ca * cb, fb + fa) for ca, fa in self.terms for cb, fb in other.terms])

    # VectorSpace (QGate.hs:64-68) and Algebra (:70-71)
    def __rmul__(self, z) -> "QGate":
        return QGate(self.n, [(complex(z) * c, f) for c, f in self.terms])

    def __add__(self, other: "QGate") -> "QGate":
        _same(self, other)
        return QGate(self.n, self.terms + other.terms)

    def __neg__(self) -> "QGate":
        return QGate(self.n, [(-c, f) for c, f in self.terms])

    def __sub__(self, other: "QGate") -> "QGate":
        return self + (-other)

    def dense(self) -> np.ndarray:
        """Materialise (tests, small n only) by applying the gate to each basis vector on the
        host through the same factor list."""
        N = 1 << self.n
        out = np.zeros((N, N), dtype=C128)
        for c, fs in self.terms:
            M = np.eye(N, dtype=C128)
            for f in fs:
                M = _factor_dense(self.n, f) @ M
            out += c * M
        return out


def _same(a: QGate, b: QGate):
    if a.n != b.n:
        raise ValueError(f"QGate {a.n} vs QGate {b.n}")  # a type error in the reference


def _factor_dense(n, f: _Factor) -> np.ndarray:
    N, k = 1 << n, len(f.qs)
    M = np.zeros((N, N), dtype=C128)
    for col in range(N):
        if any(not (col >> (n - 1 - c)) & 1 for c in f.ctrls):
            M[col, col] = 1
            continue
        sub = 0
        for j, q in enumerate(f.qs):
            sub |= ((col >> (n - 1 - q)) & 1) << (k - 1 - j)
        for row_sub in range(1 << k):
            row = col
            for j, q in enumerate(f.qs):
                b = (row_sub >> (k - 1 - j)) & 1
                row = (row & ~(1 << (n - 1 - q))) | (b << (n - 1 - q))
            M[row, col] += f.m[row_sub, sub]
    return M


# ---- constants (QGate.hs:86-118) -------------------------------------------------------------
def ident(n: int) -> QGate:
    return QGate(n)


def _g1(m) -> QGate:
    return QGate(1, [(1.0 + 0j, [_Factor([0], m)])])


def pauliX() -> QGate:
    return _g1([[0, 1], [1, 0]])


def pauliY() -> QGate:
    return _g1([[0, -1j], [1j, 0]])


def pauliZ() -> QGate:
    return _g1([[1, 0], [0, -1]])


def hadamard() -> QGate:
    return _g1((1 / math.sqrt(2)) * np.array([[1, 1], [1, -1]], dtype=C128))


def unitary_matrix(theta: float, phi: float, lam: float) -> np.ndarray:
    """QGate.hs:112-118 verbatim (NOT the OpenQASM U; not unitary in general)."""
    cis = lambda x: complex(math.cos(x), math.sin(x))
    a = cis(phi + lam / 2) * complex(math.cos(theta / 2), 0)
    b = -cis(phi - lam / 2) * complex(math.sin(theta / 2), 0)
    c = cis(phi - lam / 2) * complex(math.sin(theta / 2), 0)
    d = cis(phi + lam / 2) * complex(math.cos(theta / 2), 0)
    return np.array([[a, b], [c, d]], dtype=C128)


def unitary(theta: float, phi: float, lam: float) -> QGate:
    return _g1(unitary_matrix(theta, phi, lam))


# ---- combinators (QGate.hs:121-165) ----------------------------------------------------------
def _single_matrix(g: QGate) -> np.ndarray:
    if g.n != 1:
        raise ValueError("expected a QGate 1")
    return g.dense()


def onJust(n: int, i: int, g: QGate) -> QGate:
    """QGate.hs:148-154."""
    if not 0 <= i < n:
        raise IndexError("finite: qubit index out of range")
    return QGate(n, [(c, [f.shifted(i) for f in fs]) for c, fs in g.terms]) if g.n == 1 else _bad()


def _bad():
    raise ValueError("onJust / onEvery / onRange take a QGate 1")


def onEvery(n: int, g: QGate) -> QGate:
    """QGate.hs:158-160: n-fold Kronecker power."""
    out = ident(n)
    for i in range(n):
        out = onJust(n, i, g) @ out
    return out


def onRange(n: int, f: int, l: int, g: QGate) -> QGate:
    """QGate.hs:164-165: mconcat [onJust i m | i <- [f..l]]."""
    out = ident(n)
    for i in reversed(range(f, l + 1)):
        out = onJust(n, i, g) @ out
    return out


def kronecker(a: QGate, b: QGate) -> QGate:
    """QGate.hs:142-144: a on the first a.n qubits, b on the rest."""
    n = a.n + b.n
    wa = QGate(n, [(c, list(fs)) for c, fs in a.terms])
    wb = QGate(n, [(c, [f.shifted(a.n) for f in fs]) for c, fs in b.terms])
    return wa @ wb


def controlled(i: int, g: QGate) -> QGate:
    """QGate.hs:125-132: M.P + I - P with P = diag(bit_i).  For factors that do not touch
    qubit i this is the ordinary controlled gate (controls compose); otherwise the literal
    formula is evaluated on the few qubits involved and applied as a dense block."""
    n = g.n
    if not 0 <= i < n:
        raise IndexError("finite: qubit index out of range")
    if len(g.terms) == 1 and g.terms[0][0] == 1 and all(
            i not in f.qs and i not in f.ctrls for f in g.terms[0][1]):
        return QGate(n, [(1.0 + 0j, [_Factor(f.qs, f.m, f.ctrls + (i,)) for f in g.terms[0][1]])])
    qs = sorted({i} | {q for _, fs in g.terms for f in fs for q in f.qs + f.ctrls})
    k = len(qs)
    pos = {q: j for j, q in enumerate(qs)}
    small = QGate(k, [(c, [_Factor([pos[q] for q in f.qs], f.m, [pos[q] for q in f.ctrls]) for f in fs])
                      for c, fs in g.terms]).dense()
    j = np.arange(1 << k)
    P = np.diag(((j >> (k - 1 - pos[i])) & 1).astype(C128))
    M = small @ P + np.eye(1 << k, dtype=C128) - P
    return QGate(n, [(1.0 + 0j, [_Factor(qs, M)])])


def cnot(n: int, c: int, t: int) -> QGate:
    """QGate.hs:121-122."""
    return controlled(c, onJust(n, t, pauliX()))


def ifBit(b: int, g: QGate) -> QGate:
    """QGate.hs:136-137."""
    return g if b == 1 else ident(g.n)


# ---- application (QGate.hs:78-84) ------------------------------------------------------------
def _emit(sv: StateVec, f: _Factor):
    if len(f.qs) == 1:
        if f.ctrls:
            sv.apply_ctrl_1q(list(f.ctrls), f.qs[0], f.m)
        else:
            sv.apply_1q(f.qs[0], f.m)
    else:
        sv.apply_kq(list(f.qs), f.m, list(f.ctrls))


def gate(g: QGate, sv: StateVec) -> None:
    """QGate.hs:83-84 (StateT form): sv <- g #> sv, in place."""
    if g.n != sv.n:
        raise ValueError(f"QGate {g.n} applied to StateVec {sv.n}")
    if len(g.terms) == 1:
        c, fs = g.terms[0]
        for f in fs:
            _emit(sv, f)
        if c != 1:
            sv.scale_(c)
        return
    src = sv.clone()  # a sum of products: A v + B v + ...
    sv.scale_(0.0)
    for c, fs in g.terms:
        t = src.clone()
        for f in fs:
            _emit(t, f)
        sv.axpy_(c, t)


def apply(g: QGate, sv: StateVec) -> StateVec:
    """QGate.hs:78-80 ``g #> sv`` (pure): the argument stays valid."""
    out = sv.clone()
    gate(g, out)
    return out
