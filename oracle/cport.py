"""ctypes loader for oracle/csrc/sv_struct.c.  TEST INFRASTRUCTURE / CPU baseline only.

PARITY UNPINNED: see oracle/__init__.py.  Build with ``make -C oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SvOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("nctrl", C.c_int32),
                ("ctrl", C.c_int32 * 4), ("bit", C.c_int32), ("m", C.c_double * 8),
                ("r", C.c_double)]


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libsv_oracle.so")
    src = os.path.join(_HERE, "csrc", "sv_struct.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libsv_oracle.so"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsv_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.sv_num_threads.restype = C.c_int
        L.sv_set_threads.argtypes = [C.c_int]
        L.sv_apply_1q.argtypes = [dp, C.c_int, C.c_int, dp, C.POINTER(C.c_int), C.c_int]
        L.sv_apply_cnot.argtypes = [dp, C.c_int, C.c_int, C.c_int]
        L.sv_sumsq.argtypes = [dp, C.c_int, C.c_int, dp, dp]
        L.sv_collapse.argtypes = [dp, C.c_int, C.c_int, C.c_int]
        L.sv_measure_qubit.argtypes = [dp, C.c_int, C.c_int, C.c_double, dp]
        L.sv_measure_qubit.restype = C.c_int
        L.sv_init_basis.argtypes = [dp, C.c_int]
        L.sv_run_ops.argtypes = [dp, C.c_int, C.POINTER(SvOp), C.c_int64]
        _LIB = L
    return _LIB


def pack_ops(ops):
    """Convert an op stream (oracle.structured.run_ops format) to a C array."""
    arr = (SvOp * len(ops))()
    for o, op in zip(arr, ops):
        k = op[0]
        if k == "U":
            o.kind, o.target, o.nctrl = 0, op[1], 0
            m = np.asarray(op[2], dtype=np.complex128).reshape(4)
        elif k == "CU":
            o.kind, o.target, o.nctrl = 0, op[2], len(op[1])
            for i, c in enumerate(op[1]):
                o.ctrl[i] = c
            m = np.asarray(op[3], dtype=np.complex128).reshape(4)
        elif k == "CX":
            o.kind, o.target, o.nctrl = 1, op[2], 1
            o.ctrl[0] = op[1]
            m = None
        elif k == "COLLAPSE":
            o.kind, o.target, o.bit = 2, op[1], op[2]
            m = None
        elif k == "MEASURE":
            o.kind, o.target, o.r = 3, op[1], op[2]
            m = None
        else:
            raise ValueError(f"op {k} not supported by the C port")
        if m is not None:
            for i in range(4):
                o.m[2 * i], o.m[2 * i + 1] = m[i].real, m[i].imag
    return arr


def run_ops(n: int, ops, v: np.ndarray, packed=None) -> np.ndarray:
    """Apply an op stream in place on a copy of ``v`` with the C port; returns the copy."""
    out = np.ascontiguousarray(v, dtype=np.complex128).copy()
    arr = packed if packed is not None else pack_ops(ops)
    lib().sv_run_ops(out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), n, arr, len(arr))
    return out


def run_ops_inplace(n: int, packed, v: np.ndarray) -> None:
    assert v.dtype == np.complex128 and v.flags.c_contiguous
    lib().sv_run_ops(v.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)), n, packed, len(packed))
