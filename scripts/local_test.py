"""How fast are passes whose transposes are all warp-local?"""
import os, sys, time, re
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200 import capi
n = 30
ctx = Q.Context.default(); sv = Q.mkStateVec(n)
ctx.set_option("peephole", 0)
U = np.array([[0.8, -0.6], [0.6, 0.8]]); G = Q.unitary_matrix(.3, .2, .1)
def t(ops, reps=3):
    sv.submit(ops); sv.flush(); ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(reps): sv.submit(ops); sv.flush()
    ctx.sync(); ms = (time.perf_counter() - t0) / reps * 1e3
    st = ctx.stats()
    return ms, st["passes"] / reps, st["rounds"] / reps
for label, qs in (("9 low bits", range(n - 9, n)), ("12 low bits", range(n - 12, n)), ("bits 3..11", range(n - 12, n - 3)), ("4 hi + 5 low", list(range(0, 4)) + list(range(n - 5, n)))):
    for layers in (1, 2, 3):
        for M, name in ((U, "real"), (G, "general")):
            ops = [("U", q, M) for _ in range(layers) for q in qs]
            txt = capi.plan_describe(n, ops, "peephole=0")
            loc = [int(x) for x in re.findall(r'round [1-9]\d* regs=\[[\d,]+\] local=(\d)', txt)]
            ms, p, r = t(capi.pack_ops(ops))
            print(f"{label:14s} layers={layers} {name:8s} gates={len(ops):3d} passes={p:.0f} rounds={r:.0f} local={sum(loc)}/{len(loc)} ms={ms:.2f}", flush=True)
