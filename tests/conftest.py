import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_MAX_ERR = {}


def note_error(err, tol):
    """Parity tests report every |GPU - oracle| maximum they compare; the session writes the largest
    per test (and its tolerance) to gpurun_out/parity_max_err.json when that directory exists."""
    name = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    old = _MAX_ERR.get(name, (0.0, tol))
    _MAX_ERR[name] = (max(old[0], err), tol)


def pytest_sessionfinish(session, exitstatus):
    out = os.path.join(ROOT, "gpurun_out")
    if _MAX_ERR and os.path.isdir(out):
        import json
        worst = max(v[0] for v in _MAX_ERR.values())
        with open(os.path.join(out, "parity_max_err.json"), "w") as f:
            json.dump({"worst": worst, "tests": {k: {"max_abs_err": v[0], "tol": v[1]} for k, v in sorted(_MAX_ERR.items())}}, f, indent=1)


@pytest.fixture(scope="session")
def ctx():
    """The process-wide device context.  Fails loudly (no fallback) if the CUDA library or a
    device is missing -- GPU tests must never pass on a silent CPU path."""
    import qubism_b200 as Q
    return Q.Context.default()


@pytest.fixture()
def default_opts(ctx):
    """Restore planner options after a test that changes them."""
    names = ["tile_bits", "reg_bits", "low_bits", "max_rounds", "peephole", "fuse", "max_pass_gates", "rot", "lite", "lane_fixed", "jit", "tma", "linear", "oop", "oop_low_bits", "l2_prefetch"]
    saved = {k: ctx.get_option(k) for k in names}
    yield ctx
    for k, v in saved.items():
        ctx.set_option(k, v)


@pytest.fixture(scope="session")
def emul():
    """Test-only host emulator of k_fused_pass (tests/emul)."""
    import ctypes as C
    from qubism_b200 import capi
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d, "libqb_emul.so"], stdout=subprocess.DEVNULL)
    E = C.CDLL(os.path.join(d, "libqb_emul.so"))
    E.qbe_run.argtypes = [C.c_int, C.POINTER(capi.QbOp), C.c_int64, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64)]
    E.qbe_run.restype = C.c_int

    def run(n, ops, v, options=""):
        a = np.ascontiguousarray(v, dtype=np.complex128).copy()
        st = (C.c_int64 * 6)()
        arr = capi.pack_ops(ops)
        rc = E.qbe_run(n, arr, len(arr), options.encode(), a.ctypes.data_as(C.c_void_p), st)
        assert rc == 0, f"emulator rc={rc}"
        return a, dict(passes=st[0], rounds=st[1], gates=st[2], bank=st[3], transposes=st[4], local=st[5])

    run.lib = E
    return run


def rand_unitary_ref(rng):
    """QGateSpec.hs:14-19: unitary theta phi lambda, angles ~ U(0, 4 pi) (non-unitary)."""
    from oracle import dense
    return dense.unitary(*rng.uniform(0, 4 * np.pi, 3))
