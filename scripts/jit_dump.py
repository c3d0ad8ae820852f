"""Offline SASS check of the specialised kernels (no GPU): dump the device source of every pass of
the benchmark plan and compile one with nvcc.   python scripts/jit_dump.py OUTDIR [options] [n]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qubism_b200 import capi
from qubism_b200.circuits import qft_ops, random_layers
out = sys.argv[1]
opts = sys.argv[2] if len(sys.argv) > 2 else ""
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
os.makedirs(out, exist_ok=True)
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "libqb_emul.so"], stdout=subprocess.DEVNULL)
E = C.CDLL(os.path.join(ROOT, "tests", "emul", "libqb_emul.so"))
E.qbe_jit_dump.argtypes = [C.c_int, C.POINTER(capi.QbOp), C.c_int64, C.c_char_p, C.c_char_p, C.POINTER(C.c_int64)]
ops = capi.pack_ops(qft_ops(n) + random_layers(n, 20, seed=1000))
st = (C.c_int64 * 4)()
rc = E.qbe_jit_dump(n, ops, len(ops), opts.encode(), out.encode(), st)
print("passes", rc, "specialised", st[0], "transposes", st[1], "swizzles fixed", st[2], "left with conflicts", st[3])
