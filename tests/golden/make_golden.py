"""Generates the golden fixtures in this directory from oracle.dense (the literal restatement
of the reference's algorithm).  The Haskell reference cannot be built here (no GHC), so these
vectors pin the ORACLE, not the reference binary: "parity unpinned" (oracle/__init__.py).

    python tests/golden/make_golden.py
"""
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import dense as D  # noqa: E402
from oracle import qasm  # noqa: E402

TELEPORT = """
OPENQASM 2.0;
include "qelib1.inc";
qreg q[3];
creg c0[1]; creg c1[1]; creg c2[1];
gate post q { }
u3(0.3,0.2,0.1) q[0];
h q[1];
cx q[1],q[2];
barrier q;
cx q[0],q[1];
h q[0];
measure q[0] -> c0[0];
measure q[1] -> c1[0];
if(c0==1) z q[2];
if(c1==1) x q[2];
post q[2];
measure q[2] -> c2[0];
"""

FOURIER4 = """
OPENQASM 2.0;
include "qelib1.inc";
qreg q[4];
creg c[4];
x q[0];
x q[2];
barrier q;
h q[0];
cu1(pi/2) q[1],q[0];
h q[1];
cu1(pi/4) q[2],q[0];
cu1(pi/2) q[2],q[1];
h q[2];
cu1(pi/8) q[3],q[0];
cu1(pi/4) q[3],q[1];
cu1(pi/2) q[3],q[2];
h q[3];
measure q -> c;
"""

INVQFT4 = """
OPENQASM 2.0;
include "qelib1.inc";
qreg q[4];
creg c[4];
h q;
barrier q;
h q[0];
measure q[0] -> c[0];
if(c==1) u1(pi/2) q[1];
h q[1];
measure q[1] -> c[1];
if(c==1) u1(pi/4) q[2];
if(c==2) u1(pi/2) q[2];
if(c==3) u1(pi/2+pi/4) q[2];
h q[2];
measure q[2] -> c[2];
if(c==1) u1(pi/8) q[3];
if(c==2) u1(pi/4) q[3];
if(c==3) u1(pi/4+pi/8) q[3];
if(c==4) u1(pi/2) q[3];
if(c==5) u1(pi/2+pi/8) q[3];
if(c==6) u1(pi/2+pi/4) q[3];
if(c==7) u1(pi/2+pi/4+pi/8) q[3];
h q[3];
measure q[3] -> c[3];
"""

ADDER = """
OPENQASM 2.0;
include "qelib1.inc";
gate majority a,b,c { cx c,b; cx c,a; ccx a,b,c; }
gate unmaj a,b,c { ccx a,b,c; cx c,a; cx a,b; }
qreg cin[1];
qreg a[2];
qreg b[2];
qreg cout[1];
creg ans[3];
x a[0];
x b;
majority cin[0],b[0],a[0];
majority a[0],b[1],a[1];
cx a[1],cout[0];
unmaj a[0],b[1],a[1];
unmaj cin[0],b[0],a[0];
measure b[0] -> ans[0];
measure b[1] -> ans[1];
measure cout[0] -> ans[2];
"""

PROGRAMS = {"teleportation": (TELEPORT, 3), "fourier4": (FOURIER4, 4), "invqft4": (INVQFT4, 4), "adder2": (ADDER, 3)}


def cplx(v):
    return [[float(z.real), float(z.imag)] for z in np.asarray(v).reshape(-1)]


def serialise_trace(trace):
    out = []
    for t in trace:
        row = []
        for x in t:
            if isinstance(x, np.ndarray):
                row.append({"m": cplx(x)})
            else:
                row.append(x)
        out.append(row)
    return out


def run_program(name, forced, ref_faithful=True):
    src, nmeas = PROGRAMS[name]
    trace = []
    ps = qasm.run_qasm(src, backend=qasm.DenseBackend(), draws=list(forced), ref_faithful=ref_faithful, trace=trace)
    states = {k: cplx(v[1]) for k, v in ps.stVecs.items()}
    # A forced outcome whose weight is only rounding noise (an impossible branch) amplifies that
    # noise to a unit vector: such paths are ill-conditioned and are flagged, not compared.
    degenerate = False
    for t in trace:
        if t[0] == "MEASURE":
            bit, p = t[4], t[5]
            w = (p * p if bit == 1 else 1.0 - p * p) if p == p else (1.0 if bit == 0 else 0.0)
            degenerate = degenerate or w < 1e-9 or (p == p and abs(t[3] - p) < 1e-9)  # or a knife-edge draw
    return {"draws": list(forced), "cregs": ps.cregs, "states": states, "trace": serialise_trace(trace),
            "degenerate": degenerate}


def main():
    gold = {}
    for name, (src, nmeas) in PROGRAMS.items():
        runs = []
        for forced in itertools.product([2.0, -1.0], repeat=nmeas):  # 2.0 forces Zero, -1.0 forces One
            runs.append(run_program(name, forced))
        runs.append(run_program(name, (0.37,) * nmeas))  # the rule decides: r < pOne
        gold[name] = {"source": src, "runs": runs}
    # adder once more with the write-back bug repaired, to show what the bug changes
    gold["adder2_fixed"] = {"source": ADDER, "runs": [run_program("adder2", (0.37, 0.37, 0.37), ref_faithful=False)]}
    # raw gate-level vectors on random (reference-distribution) states
    rng = np.random.default_rng(20181018)
    cases = []
    for n in (1, 2, 3, 5):
        v = rng.uniform(-1, 1, 1 << n) + 1j * rng.uniform(-1, 1, 1 << n)
        v = v / np.linalg.norm(v)
        for q in range(n):
            th, ph, la = rng.uniform(0, 4 * np.pi, 3)
            m = D.unitary(th, ph, la)
            cases.append({"n": n, "op": ["U", q, {"angles": [th, ph, la]}], "in": cplx(v),
                          "out": cplx(D.apply(D.onJust(n, q, m), v))})
            for b in (0, 1):
                with np.errstate(all="ignore"):
                    cases.append({"n": n, "op": ["COLLAPSE", q, b], "in": cplx(v), "out": cplx(D.collapse(n, q, b, v))})
            for c in range(n):
                if c != q:
                    cases.append({"n": n, "op": ["CX", c, q], "in": cplx(v), "out": cplx(D.apply(D.cnot(n, c, q), v))})
                    cases.append({"n": n, "op": ["CU", [c], q, {"angles": [th, ph, la]}], "in": cplx(v),
                                  "out": cplx(D.apply(D.controlled(n, c, D.onJust(n, q, m)), v))})
    gold["gate_vectors"] = cases
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f)
    print("wrote golden.json:", {k: (len(v["runs"]) if isinstance(v, dict) else len(v)) for k, v in gold.items()})


if __name__ == "__main__":
    main()
