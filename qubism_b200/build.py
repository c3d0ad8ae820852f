"""Builds libqubism_sv.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m qubism_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libqubism_sv.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU = ["qb_kernels.cu"]
CPP = ["qb_api.cpp", "qb_planner.cpp", "qb_dist.cpp", "qb_jit.cpp"]
HDRS = ["qb_internal.h", "qb_kernels.h", "qb_dist.h", "qb_jit.h", "qb_jit_prelude.cuh",
        os.path.join("..", "..", "include", "qubism_sv.h")]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def embed_prelude(obj_dir: str = OBJ) -> str:
    """qb_jit_prelude.cuh as a C++ raw string literal (qb_jit.cpp #includes it): the source of the
    structure-specialised kernels travels inside the library."""
    os.makedirs(obj_dir, exist_ok=True)
    src = os.path.join(CSRC, "qb_jit_prelude.cuh")
    dst = os.path.join(obj_dir, "qb_jit_prelude.inc")
    text = 'R"QBJIT(' + open(src).read() + ')QBJIT"\n'
    if not os.path.exists(dst) or open(dst).read() != text:
        with open(dst, "w") as f:
            f.write(text)
    return dst


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    embed_prelude()
    hdrs = [os.path.join(CSRC, h) for h in HDRS] + [os.path.abspath(__file__)]
    objs = []
    for src in CU + CPP:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src + ".o")
        objs.append(op)
        if not (force or _stale(op, [sp] + hdrs)):
            continue
        cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall", "-I", OBJ, "-c", sp, "-o", op]
        if src.endswith(".cu"):
            cmd += ["-Xptxas", "-v"] if verbose else []
        else:
            cmd += ["-x", "cu"] if False else []
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
