"""Literal dense restatement of the reference's gate/state algebra.  TEST INFRASTRUCTURE.

Every function follows one reference definition and cites it (paths relative to
/root/reference).  This is the reference's *real* algorithm: every gate is a dense
2^n x 2^n complex matrix built with Kronecker products, ``controlled`` is a dense
matrix product, application is a dense mat-vec.  Usable for n <= 12 or so
(SURVEY.md section 6).  hmatrix semantics relied upon are listed in SURVEY.md 8(c).

PARITY UNPINNED: see oracle/__init__.py.

Conventions: a state is a 1-D complex128 array of length 2^n, a gate is a
(2^n, 2^n) complex128 array, qubit i is bit n-1-i of the amplitude index
(StateVec.hs:65-67), Bit is the int 0 (Zero) / 1 (One) (CReg.hs:14).
"""
from __future__ import annotations

import math

import numpy as np

C = np.complex128  # Algebra.hs:14  type C = Complex Double


# --------------------------------------------------------------------------- gates
def ident(n: int) -> np.ndarray:
    """QGate.hs:86-87  ident = LA.ident (2^n)."""
    return np.eye(1 << n, dtype=C)


def pauliX() -> np.ndarray:
    """QGate.hs:90-93."""
    return np.array([[0, 1], [1, 0]], dtype=C)


def pauliY() -> np.ndarray:
    """QGate.hs:95-98."""
    return np.array([[0, -1j], [1j, 0]], dtype=C)


def pauliZ() -> np.ndarray:
    """QGate.hs:100-103."""
    return np.array([[1, 0], [0, -1]], dtype=C)


def hadamard() -> np.ndarray:
    """QGate.hs:105-108  1 / sqrt 2 * (2><2)[1,1,1,-1]."""
    return (1 / math.sqrt(2)) * np.array([[1, 1], [1, -1]], dtype=C)


def cis(x: float) -> complex:
    """Data.Complex.cis."""
    return complex(math.cos(x), math.sin(x))


def unitary(theta: float, phi: float, lam: float) -> np.ndarray:
    """QGate.hs:112-118.  NOT the OpenQASM U and not unitary in general."""
    a = cis(phi + lam / 2) * complex(math.cos(theta / 2), 0)
    b = -cis(phi - lam / 2) * complex(math.sin(theta / 2), 0)
    c = cis(phi - lam / 2) * complex(math.sin(theta / 2), 0)
    d = cis(phi + lam / 2) * complex(math.cos(theta / 2), 0)
    return np.array([[a, b], [c, d]], dtype=C)


def kronecker(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """QGate.hs:142-144  LA.kronecker a b (a acts on the first qubits)."""
    return np.kron(a, b)


def onJust(n: int, i: int, m: np.ndarray) -> np.ndarray:
    """QGate.hs:148-154  ident(2^i) (x) m (x) ident(2^(n-i-1))."""
    if not 0 <= i < n:
        raise IndexError("finite: qubit index out of range")  # Data.Finite `finite`
    pre = np.eye(1 << i, dtype=C)
    post = np.eye(1 << (n - i - 1), dtype=C)
    return np.kron(np.kron(pre, m), post)


def onEvery(n: int, m: np.ndarray) -> np.ndarray:
    """QGate.hs:158-160  iterate (kronecker m) m !! (n-1)."""
    g = m
    for _ in range(n - 1):
        g = np.kron(m, g)
    return g


def mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """QGate.hs:58-59  (<>) = matrix product."""
    return a @ b


def onRange(n: int, f: int, l: int, m: np.ndarray) -> np.ndarray:
    """QGate.hs:164-165  mconcat [onJust i m | i <- [f..l]]  (foldr (<>) ident)."""
    g = ident(n)
    for i in reversed(range(f, l + 1)):
        g = onJust(n, i, m) @ g
    return g


def controlled(n: int, i: int, m: np.ndarray) -> np.ndarray:
    """QGate.hs:125-132  (m <> projection) + ident - projection,
    projection = diag [ j `quot` 2^(n-i-1) `mod` 2 | j <- [0..] ]."""
    if not 0 <= i < n:
        raise IndexError("finite: qubit index out of range")
    j = np.arange(1 << n)
    projection = np.diag(((j // (1 << (n - i - 1))) % 2).astype(C))
    return (m @ projection) + np.eye(1 << n, dtype=C) - projection


def cnot(n: int, c: int, t: int) -> np.ndarray:
    """QGate.hs:121-122  controlled c . onJust t $ pauliX."""
    return controlled(n, c, onJust(n, t, pauliX()))


def ifBit(n: int, b: int, g: np.ndarray) -> np.ndarray:
    """QGate.hs:136-137."""
    return g if b == 1 else ident(n)


def apply(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """QGate.hs:78-80  (#>) = dense mat-vec."""
    return g @ v


# -------------------------------------------------------------------------- states
def mkStateVec(n: int) -> np.ndarray:
    """StateVec.hs:78-85  |0...0>."""
    v = np.zeros(1 << n, dtype=C)
    v[0] = 1
    return v


def zero(n: int) -> np.ndarray:
    """StateVec.hs:52."""
    return np.zeros(1 << n, dtype=C)


def normalize(v: np.ndarray) -> np.ndarray:
    """StateVec.hs:91-92  LA.normalize = v / norm_2 v  (0/0 -> NaN, as the reference)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return v / np.linalg.norm(v)


def tensor(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """StateVec.hs:98-100  flatten (outer a b): out[i*2^m + j] = a_i * b_j, no conjugation."""
    return np.outer(a, b).reshape(-1)


def inner(a: np.ndarray, b: np.ndarray) -> complex:
    """StateVec.hs:57-58  LA.<.> conjugates the FIRST argument (AlgebraTests.hs:43-47)."""
    return complex(np.vdot(a, b))


def norm(a: np.ndarray) -> float:
    """Algebra.hs:35-36  norm a = realPart (a <.> a)  -- the SQUARED 2-norm."""
    return inner(a, a).real


def approx_eq(a: np.ndarray, b: np.ndarray) -> bool:
    """StateVec.hs:47-49 / QGate.hs:54-56  norm_2 (a - b) < 1e-6."""
    return bool(np.linalg.norm(a - b, 2) < 0.000001)


def collapse(n: int, i: int, b: int, v: np.ndarray) -> np.ndarray:
    """StateVec.hs:104-114  normalize (v * mask), mask = blocks of m = 2^n / 2^(i+1)
    entries alternating ifZero, ifOne."""
    l = 1 << n
    m = l // (1 << (i + 1))
    ifZero = 1.0 if b == 0 else 0.0
    ifOne = 1.0 if b == 1 else 0.0
    mask = np.tile(np.concatenate([np.full(m, ifZero), np.full(m, ifOne)]), l // (2 * m))
    return normalize(v * mask.astype(C))


def measureQubit(n: int, i: int, r: float, v: np.ndarray):
    """StateVec.hs:118-129 with the uniform draw ``r`` supplied by the caller.
    Returns (bit, new_state, pOne).  pOne = Re <collapse i One v | v> = sqrt(S1); when
    S1 == 0 it is NaN in the reference and ``r < NaN`` is False -> Zero."""
    qrZero = collapse(n, i, 0, v)
    qrOne = collapse(n, i, 1, v)
    with np.errstate(invalid="ignore"):
        pOne = inner(qrOne, v).real
    if r < pOne:
        return 1, qrOne, pOne
    return 0, qrZero, pOne


def measure(n: int, rs, v: np.ndarray):
    """StateVec.hs:133-137  traverse measureQubit [0..n-1]; bit i of the CReg = qubit i."""
    bits = []
    for i in range(n):
        b, v, _ = measureQubit(n, i, rs[i], v)
        bits.append(b)
    return bits, v


def crToNatural(bits) -> int:
    """CReg.hs:36-39  little-endian: bit i weighs 2^i."""
    return sum((1 << i) for i, b in enumerate(bits) if b == 1)


def show(n: int, v: np.ndarray) -> str:
    """StateVec.hs:60-68  '% 6.4f  + % 6.4fi  |bits>' per amplitude."""
    out = []
    for i, z in enumerate(v):
        bits = "".join("0" if (i // (1 << (n - j - 1))) % 2 == 0 else "1" for j in range(n))
        out.append("% 6.4f" % z.real + "  + " + "% 6.4f" % z.imag + "i" + "  " + "|" + bits + ">\n")
    return "".join(out)
