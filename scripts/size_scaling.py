"""Is a gate-heavy pass compute- or memory-limited?  Same 24-gate / 4-round program on the 12 low
bits at n = 22 (state fits in L2) ... 30 (HBM): time per amplitude should be flat if compute-bound."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200 import capi
ctx = Q.Context.default()
ctx.set_option("peephole", 0)
U = np.array([[0.8, -0.6], [0.6, 0.8]])
for n in (22, 24, 26, 28, 30):
    sv = Q.mkStateVec(n)
    for label, layers in (("1 gate", 0), ("12 real", 1), ("24 real", 2), ("48 real", 4)):
        ops = [("U", n - 1, U)] if layers == 0 else [("U", q, U) for _ in range(layers) for q in range(n - 12, n)]
        arr = capi.pack_ops(ops)
        reps = max(3, 1 << max(0, 28 - n))
        sv.submit(arr); sv.flush(); ctx.sync(); ctx.reset_stats()
        t0 = time.perf_counter()
        for _ in range(reps): sv.submit(arr); sv.flush()
        ctx.sync(); ms = (time.perf_counter() - t0) / reps * 1e3
        st = ctx.stats()
        print(f"n={n} {label:8s} passes={st['passes']/reps:.0f} rounds={st['rounds']/reps:.0f} ms={ms:8.3f}  ns/amp={ms*1e6/(1<<n):.4f}  scaled-to-30q ms={ms*(1<<(30-n)):.2f}", flush=True)
    del sv
