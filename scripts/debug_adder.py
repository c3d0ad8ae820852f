import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import qubism_b200 as Q
from oracle import qasm
from test_gpu_parity import GpuBackend
G = json.load(open("tests/golden/golden.json"))
src = G["adder2"]["source"]
ast = qasm.parse(src)
draws = [0.37] * 3
def mk(be):
    it = iter(draws)
    return qasm.Evaluator(be, lambda: next(it), ref_faithful=True, trace=[])
eg, eo = mk(GpuBackend()), mk(qasm.StructuredBackend())
def flat(stmts):
    for s in stmts:
        if s[0] == "PosInfo" and s[2][0] == "StmtList":
            yield s  # include expands as a unit
        else:
            yield s
for s in ast:
    ng, no = len(eg.trace), len(eo.trace)
    eg.run_stmt(s); eo.run_stmt(s)
    keys = sorted(set(eg.ps.stVecs) | set(eo.ps.stVecs))
    bad = []
    for k in keys:
        if k not in eg.ps.stVecs or k not in eo.ps.stVecs:
            bad.append((k, "missing")); continue
        a = eg.ps.stVecs[k].to_host(); b = eo.ps.stVecs[k][1]
        d = np.abs(a - b).max() if a.shape == b.shape else -1
        if not d < 1e-12: bad.append((k, float(d)))
    if bad:
        print("line", s[1], s[2][0], "new trace:", [(t[0],) + tuple(x for x in t[1:4] if not isinstance(x, np.ndarray)) for t in eo.trace[no:]][:6], "MISMATCH", bad)
        break
else:
    print("all statements match", eg.ps.cregs, eo.ps.cregs)
