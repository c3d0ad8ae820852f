"""CPU-only tests of the structure-specialised fused passes (qubism_b200/csrc/qb_jit.cpp).

The generator emits, for one planned step pass, either the CUDA source NVRTC compiles on the GPU
box or a C++ emulation of the same code (threads of a CTA run phase by phase).  Here the
emulation is compiled with g++ and checked against the oracle; the CUDA flavour is compiled to a
cubin with NVRTC (which needs no device).  Parity of the real kernels is in test_gpu_parity.py.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import dense as D, structured as S  # noqa: E402
from qubism_b200 import capi  # noqa: E402
from qubism_b200.circuits import qft_ops, random_layers, random_mixed  # noqa: E402


@pytest.fixture(scope="session")
def jit_emul(tmp_path_factory):
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d, "libqb_emul.so"], stdout=subprocess.DEVNULL)
    E = C.CDLL(os.path.join(d, "libqb_emul.so"))
    E.qbe_run_jit.argtypes = [C.c_int, C.POINTER(capi.QbOp), C.c_int64, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64),
                              C.c_char_p, C.c_int, C.c_char_p, C.c_int64]
    E.qbe_run_jit.restype = C.c_int
    work = str(tmp_path_factory.mktemp("qbj"))

    def run(n, ops, v, options="", src_index=-1):
        a = np.ascontiguousarray(v, dtype=np.complex128).copy()
        st = (C.c_int64 * 6)()
        arr = capi.pack_ops(ops)
        cap = 1 << 21
        buf = C.create_string_buffer(cap) if src_index >= 0 else None
        rc = E.qbe_run_jit(n, arr, len(arr), options.encode(), a.ctypes.data_as(C.c_void_p), st, work.encode(),
                           src_index, buf, cap)
        assert rc == 0, f"qbe_run_jit rc={rc} (see {work}/*.log)"
        return a, dict(passes=st[0], jit=st[1], structures=st[2], guards=st[3]), (buf.value.decode() if buf else None)

    return run


@pytest.mark.parametrize("opts", ["", "reg_bits=3", "reg_bits=5", "tile_bits=10,reg_bits=4", "low_bits=5", "rot=0",
                                  "lane_fixed=1", "lane_fixed=2,reg_bits=5", "max_pass_gates=7", "tile_bits=11,reg_bits=3",
                                  "tile_bits=13,reg_bits=4", "tile_bits=13,reg_bits=5,low_bits=6", "max_rounds=3", "lane_fixed=3",
                                  "tma=1", "tma=1,reg_bits=5", "tma=1,tile_bits=11,reg_bits=3", "tma=1,lane_fixed=3",
                                  "tma=1,max_pass_gates=5", "tma=1,low_bits=5", "tma=2", "tma=2,reg_bits=5",
                                  "tma=2,tile_bits=11,reg_bits=3", "tma=2,max_pass_gates=4"])
def test_specialised_passes_of_rotation_cx_layers_match_oracle(jit_emul, opts):
    """U(theta, phi, 0) + CX layers: every pass is a step pass, every one is specialised: 2-FMA
    rotations with deferred cosines (forms A and B), flip-aware flavours after toggles, static
    and masked register swaps, folded flips in transposes and stores."""
    n = 13
    ops = random_layers(n, 6, seed=77, lam0=True)
    v = S.gen_state(n, np.random.default_rng(5))
    ref = S.run_ops(n, ops, v)
    out, st, _ = jit_emul(n, ops, v, opts)
    assert st["jit"] == st["passes"] > 0
    assert np.abs(out - ref).max() < 1e-13


@pytest.mark.parametrize("seed", range(10))
def test_specialised_passes_random_mixes(jit_emul, seed):
    """General / real / scaled-general slots, uncontrolled X (a renaming), controlled gates and
    diagonal gates (those passes stay with the generic kernels), known support."""
    rng = np.random.default_rng(4000 + seed)
    n = int(rng.integers(12, 15))
    ops = []
    for _ in range(int(rng.integers(30, 80))):
        kind = int(rng.integers(0, 10))
        q = int(rng.integers(0, n))
        if kind <= 2:
            ops.append(("U", q, D.unitary(float(rng.uniform(0, 12)), float(rng.uniform(0, 12)), 0.0)))
        elif kind == 3:
            ops.append(("U", q, D.unitary(*[float(x) for x in rng.uniform(0, 12, 3)])))
        elif kind == 4:
            ops.append(("U", q, [D.hadamard(), D.pauliX(), D.pauliY(), np.array([[1.0, 2.0], [0.5, -1.0]])][int(rng.integers(0, 4))]))
        elif kind <= 8:
            c = int(rng.integers(0, n))
            if c != q:
                ops.append(("CX", c, q))
        elif seed % 2:
            c = int(rng.integers(0, n))
            if c != q:
                ops.append(("CU", [c], q, D.unitary(*[float(x) for x in rng.uniform(0, 12, 3)])))
    knobs = []
    if rng.integers(0, 2): knobs.append(f"reg_bits={int(rng.choice([3, 4, 5]))}")
    if rng.integers(0, 2): knobs.append(f"lane_fixed={int(rng.integers(0, 3))}")
    if rng.integers(0, 2): knobs.append(f"low_bits={int(rng.integers(3, 6))}")
    if rng.integers(0, 3) == 0: knobs.append(f"max_pass_gates={int(rng.integers(4, 20))}")
    if rng.integers(0, 2): knobs.append("tma=1")
    v = S.gen_state(n, rng)
    if rng.integers(0, 2):
        qa = int(rng.integers(0, n))
        ba = int(rng.integers(0, 2))
        v = S.collapse(n, qa, ba, v)
        knobs += [f"known_mask={1 << (n - 1 - qa)}", f"known_val={ba << (n - 1 - qa)}"]
    ref = S.run_ops(n, ops, v)
    out, st, _ = jit_emul(n, ops, v, ",".join(knobs))
    if seed % 2 == 0:  # (with controlled-U gates in the mix a circuit may hold no step pass at all)
        assert st["jit"] > 0
    assert np.abs(out - ref).max() < 1e-12, knobs


def test_specialised_qft_reference_semantics(jit_emul):
    """The reference's QFT (u1 = scalar, merged h gates of the general class) and a dense run of
    layers on a fresh |0...0> with everything known."""
    n = 12
    v = S.gen_state(n, np.random.default_rng(1))
    ops = qft_ops(n) + random_layers(n, 2, seed=3, lam0=True)
    out, st, _ = jit_emul(n, ops, v)
    assert st["jit"] > 0
    assert np.abs(out - S.run_ops(n, ops, v)).max() < 1e-13
    z = np.zeros(1 << n, complex)
    z[0] = 1
    ops = random_layers(n, 3, seed=8, lam0=True)
    out, st, _ = jit_emul(n, ops, z, f"known_mask={(1 << n) - 1},known_val=0")
    assert np.abs(out - S.run_ops(n, ops, z)).max() < 1e-13


def test_equal_structure_gives_one_kernel_and_new_angles_only_new_coefficients(jit_emul):
    """The structural key holds no gate coefficient except the form (A / B) of each rotation: the
    same layers with angles nudged inside their octant reuse the compiled code."""
    n = 12
    ops = random_layers(n, 2, seed=21, lam0=True)
    v = S.gen_state(n, np.random.default_rng(2))
    _, st, src_a = jit_emul(n, ops, v, "", 0)
    ops2 = []
    for op in ops:
        if op[0] == "U":
            m = np.asarray(op[2])
            eps = 1e-3
            rot = np.array([[np.cos(eps), -np.sin(eps)], [np.sin(eps), np.cos(eps)]])
            ops2.append(("U", op[1], rot @ m))
        else:
            ops2.append(op)
    out, st2, src_b = jit_emul(n, ops2, v, "", 0)
    assert np.abs(out - S.run_ops(n, ops2, v)).max() < 1e-13
    assert src_a == src_b, "angles leaked into the generated source"


@pytest.mark.parametrize("opts", ["", "reg_bits=5", "tile_bits=10,reg_bits=3", "rot=0", "jit_group=4,jit_pf_last=0",
                                  "jit_mem=5", "jit_mem=6,jit_minb=3", "tile_bits=13,reg_bits=4", "l2_prefetch=0",
                                  "tma=1", "tma=1,reg_bits=5", "tma=1,l2_prefetch=0", "tma=1,max_pass_gates=3",
                                  "tma=2", "tma=2,reg_bits=5", "tma=2,tile_bits=11,reg_bits=3", "tma=2,max_pass_gates=3",
                                  "tma=2,tile_bits=13,reg_bits=4"])
def test_device_source_compiles_with_nvrtc(jit_emul, opts):
    """NVRTC needs no GPU: the CUDA flavour of the generated source must compile for sm_100a."""
    n = 13
    ops = random_layers(n, 4, seed=5, lam0=True) + [("U", 3, D.unitary(0.3, 0.2, 0.1)), ("U", 1, D.pauliX())]
    v = S.gen_state(n, np.random.default_rng(0))
    _, st, src = jit_emul(n, ops, v, opts, 0)
    assert src and "qb_jit_pass" in src and "__launch_bounds__" in src
    if "tma=" in opts:
        assert "qbj_bulk_load(" in src.split("qb_jit_pass(")[1], "the bulk-asynchronous load path was not generated"
    nbytes = C.c_int64(0)
    rc = capi.lib().qb_jit_compile_check(src.encode(), C.byref(nbytes))
    if rc == capi.QB_ERR_UNSUPPORTED:
        pytest.skip("libnvrtc not available: " + capi.lib().qb_last_error().decode())
    assert rc == 0, capi.lib().qb_last_error().decode()[:2000]
    assert nbytes.value > 1000


@pytest.mark.parametrize("opts,lbits", [("", (11, 10)), ("", (12, 3)), ("reg_bits=5", (9, 1)), ("tile_bits=10,reg_bits=3", (12, 11))])
def test_swap_carrying_pass_compiles_with_nvrtc(jit_emul, opts, lbits):
    """The generated kernel whose stores carry a global<->local swap (option fuse_exchange): peer table
    and geometry are kernel arguments (QbjXch); the run-time bit positions are looked at once per tile
    (QBJ_XCH_TILE), a store XORs its register index's share of the rank (dr[i]) and its offset outside
    the victims (QBJ_DSTX); one flag of the structure key tells it from the plain pass."""
    n = 13
    E = C.CDLL(os.path.join(ROOT, "tests", "emul", "libqb_emul.so"))
    E.qbe_xch_source.argtypes = [C.c_int, C.POINTER(capi.QbOp), C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int64]
    ops = capi.pack_ops(random_layers(n, 4, seed=5, lam0=True))
    cap = 1 << 21
    buf = C.create_string_buffer(cap)
    assert E.qbe_xch_source(n, ops, len(ops), opts.encode(), lbits[0], lbits[1], buf, cap) == 0
    src = buf.value.decode()
    body = src.split("qb_jit_pass(")[1]
    assert "QBJ_XCH_TILE(at_)" in body and ("QBJ_ST2X(" in body or "QBJ_ST1X(" in body)
    assert "QBJ_ST2(" not in body and "QBJ_ST1(" not in body  # (no store bypasses the destination arithmetic)
    nbytes = C.c_int64(0)
    rc = capi.lib().qb_jit_compile_check(src.encode(), C.byref(nbytes))
    if rc == capi.QB_ERR_UNSUPPORTED:
        pytest.skip("libnvrtc not available: " + capi.lib().qb_last_error().decode())
    assert rc == 0, capi.lib().qb_last_error().decode()[:2000]
    assert nbytes.value > 1000


def test_range_guard_applies_the_running_factor_mid_flush(jit_emul):
    """~1,100 rotations near 90 degrees in one flush: the product of the factors their 2-FMA forms
    leave out drops below 2^-300, so a pass in the middle must apply the running factor (its
    kernel is the has_gscale flavour) and the bookkeeping restarts at 1."""
    n = 12
    rng = np.random.default_rng(314)
    ops = []
    for layer in range(90):
        for q in range(n):
            ops.append(("U", q, D.unitary(float(rng.uniform(1.2, 1.9)), float(rng.uniform(0, 6)), 0.0)))
        perm = rng.permutation(n)
        for k in range(0, n - 1, 2):
            ops.append(("CX", int(perm[k]), int(perm[k + 1])))
    v = S.gen_state(n, rng)
    out, st, _ = jit_emul(n, ops, v)
    assert st["guards"] >= 1, "the circuit was meant to trip the range guard"
    assert np.abs(out - S.run_ops(n, ops, v)).max() < 1e-12
