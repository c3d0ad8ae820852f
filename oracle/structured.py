"""Structured O(2^n)-per-op numpy restatement of the same semantics.  TEST INFRASTRUCTURE.

Same arithmetic meaning as ``oracle.dense`` (which follows the reference literally) but
without materialising 2^n x 2^n matrices, so it reaches n ~ 26 on the host.  It is
cross-checked against ``oracle.dense`` for n <= 10 in tests/test_oracle.py; SURVEY.md
Appendix A explains why the two (and the reference) agree only to ~1 ulp per op and why
parity is 1e-12 absolute, not bit-exact.

PARITY UNPINNED: see oracle/__init__.py.

Qubit i <-> axis i of ``v.reshape([2]*n)`` (big-endian, StateVec.hs:65-67).
Functions mutate nothing; they return new arrays.
"""
from __future__ import annotations

import math

import numpy as np

C = np.complex128


def _nd(v: np.ndarray, n: int) -> np.ndarray:
    return v.reshape([2] * n) if n > 0 else v.reshape(())


def apply_1q(n: int, q: int, m, v: np.ndarray, ctrls=()) -> np.ndarray:
    """``controlled c1 (.. (onJust q m)) #> v`` (QGate.hs:125-132,148-154,78-80): act with the
    2x2 ``m`` (row-major [[a,b],[c,d]]) on qubit ``q`` of the components whose control
    qubits are all 1; identity elsewhere."""
    m = np.asarray(m, dtype=C).reshape(2, 2)
    if not 0 <= q < n or any(not 0 <= c < n for c in ctrls):
        raise IndexError("finite: qubit index out of range")
    if q in ctrls:
        raise ValueError("target is also a control")
    out = v.astype(C, copy=True)
    nd = _nd(out, n)
    sel = [slice(None)] * n
    for c in ctrls:
        sel[c] = 1
    sub = nd[tuple(sel)]  # view; axes = qubits not in ctrls, in order
    ax = q - sum(1 for c in set(ctrls) if c < q)
    a0 = np.take(sub, 0, axis=ax).copy()
    a1 = np.take(sub, 1, axis=ax).copy()
    idx0 = [slice(None)] * sub.ndim
    idx1 = [slice(None)] * sub.ndim
    idx0[ax] = 0
    idx1[ax] = 1
    sub[tuple(idx0)] = m[0, 0] * a0 + m[0, 1] * a1
    sub[tuple(idx1)] = m[1, 0] * a0 + m[1, 1] * a1
    return out


def apply_cnot(n: int, c: int, t: int, v: np.ndarray) -> np.ndarray:
    """``cnot c t #> v`` (QGate.hs:121-122): swap v_k <-> v_{k xor bit_t} where bit_c = 1."""
    if c == t:
        raise ValueError("cnot with c == t")
    return apply_1q(n, t, [[0, 1], [1, 0]], v, ctrls=(c,))


def apply_kq(n: int, qs, M, v: np.ndarray, ctrls=()) -> np.ndarray:
    """Dense k-qubit block on qubits ``qs`` (qs[0] = most significant index bit of M, the
    order ``kronecker a b`` gives, QGate.hs:142-144), optionally controlled."""
    k = len(qs)
    M = np.asarray(M, dtype=C).reshape(1 << k, 1 << k)
    if len(set(qs)) != k or set(qs) & set(ctrls):
        raise ValueError("repeated qubit")
    if any(not 0 <= q < n for q in list(qs) + list(ctrls)):
        raise IndexError("finite: qubit index out of range")
    out = v.astype(C, copy=True)
    nd = _nd(out, n)
    sel = [slice(None)] * n
    for c in ctrls:
        sel[c] = 1
    sub = nd[tuple(sel)]
    axes = [q - sum(1 for c in set(ctrls) if c < q) for q in qs]
    moved = np.moveaxis(sub, axes, list(range(k)))
    shp = moved.shape
    res = (M @ moved.reshape(1 << k, -1)).reshape(shp)
    moved[...] = res  # writes through the view into ``out``
    return out


def sumsq(n: int, q: int, v: np.ndarray):
    """(S0, S1) = sum |z_k|^2 over bit_q(k) = 0 / 1.  The reference's decision value is
    pOne = sqrt(S1) (StateVec.hs:124-126 with collapse's normalisation at :107)."""
    nd = v.reshape(1 << q, 2, -1)
    a = nd.real ** 2 + nd.imag ** 2
    return float(a[:, 0, :].sum()), float(a[:, 1, :].sum())


def collapse(n: int, q: int, b: int, v: np.ndarray) -> np.ndarray:
    """StateVec.hs:104-114: mask then divide by the 2-norm (zero weight -> NaN everywhere)."""
    out = v.astype(C, copy=True)
    nd = out.reshape(1 << q, 2, -1)
    nd[:, 1 - b, :] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        return out / np.linalg.norm(out)


def measure_qubit(n: int, q: int, r: float, v: np.ndarray):
    """StateVec.hs:118-129 with the draw ``r`` supplied: One iff r < sqrt(S1) (NaN -> Zero)."""
    _, s1 = sumsq(n, q, v)
    p_one = math.sqrt(s1) if s1 > 0 else float("nan")
    if r < p_one:
        return 1, collapse(n, q, 1, v), p_one
    return 0, collapse(n, q, 0, v), p_one


def tensor(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """StateVec.hs:98-100."""
    return np.outer(a, b).reshape(-1)


def mk_state(n: int) -> np.ndarray:
    """StateVec.hs:78-85."""
    v = np.zeros(1 << n, dtype=C)
    v[0] = 1
    return v


def gen_state(n: int, rng: np.random.Generator, normalized: bool = True) -> np.ndarray:
    """test/Qubism/StateVecSpec.hs:20-28: re, im ~ U(-1, 1), then normalize."""
    v = rng.uniform(-1, 1, 1 << n) + 1j * rng.uniform(-1, 1, 1 << n)
    v = v.astype(C)
    return v / np.linalg.norm(v) if normalized else v


# ------------------------------------------------------------------ op streams
# The primitive op stream that reaches the hot path:
#   ("U", q, m2x2)                 onJust q m #>
#   ("CX", c, t)                   cnot c t #>
#   ("CU", ctrls, t, m2x2)         controlled c1 (controlled c2 ... (onJust t m)) #>
#   ("KQ", qs, M, ctrls)           dense block
#   ("COLLAPSE", q, b)             collapse q b
#   ("MEASURE", q, r)              measureQubit q with draw r
def run_ops(n: int, ops, v: np.ndarray, record=None) -> np.ndarray:
    """Apply an op stream one primitive op at a time, exactly as the reference evaluator
    does (QASM/Simulation.hs:94-122 issues one ``#>`` per primitive op)."""
    for op in ops:
        k = op[0]
        if k == "U":
            v = apply_1q(n, op[1], op[2], v)
        elif k == "CX":
            v = apply_cnot(n, op[1], op[2], v)
        elif k == "CU":
            v = apply_1q(n, op[2], op[3], v, ctrls=tuple(op[1]))
        elif k == "KQ":
            v = apply_kq(n, op[1], op[2], v, ctrls=tuple(op[3]) if len(op) > 3 else ())
        elif k == "COLLAPSE":
            v = collapse(n, op[1], op[2], v)
        elif k == "MEASURE":
            bit, v, p = measure_qubit(n, op[1], op[2], v)
            if record is not None:
                record.append((op[1], bit, p))
        else:
            raise ValueError(f"unknown op {k}")
    return v
