"""Measurement throughput: `measure` on every qubit of a dense n-qubit
state, a mid-circuit measure + continue, and a circuit started from a fresh |0...0>.
   python scripts/measure_bench.py [n]        (QB_SUPPORT=0 for the baseline without support tracking)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import random_layers
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ctx = Q.Context.default()
layer = capi.pack_ops(random_layers(n, 2, seed=1000))
rng = np.random.default_rng(1)
def timed(f):
    ctx.sync(); ctx.reset_stats(); t0 = time.perf_counter(); r = f(); ctx.sync()
    return (time.perf_counter() - t0) * 1e3, r, ctx.stats()
out = {"n": n, "support": ctx.get_option("support")}
sv = Q.mkStateVec(n)
ms, _, st = timed(lambda: (sv.submit(layer), sv.flush()))
out["two_layers_from_fresh_state_ms"] = round(ms, 2); out["tiles_fresh"] = st["tiles"]; out["passes_fresh"] = st["passes"]
ms, _, st = timed(lambda: (sv.submit(layer), sv.flush()))
out["two_layers_dense_ms"] = round(ms, 2); out["tiles_dense"] = st["tiles"]
rs = list(rng.uniform(0, 1, n))
ms, bits, st = timed(lambda: Q.measure(sv, rs))
out["measure_all_ms"] = round(ms, 2); out["measure_all_launches"] = st["reduce_launches"] + st["simple_launches"] + st["passes"]
ms, _, st = timed(lambda: (sv.submit(layer), sv.flush()))
out["two_layers_after_measure_all_ms"] = round(ms, 2); out["tiles_after_measure"] = st["tiles"]
# mid-circuit: measure 4 qubits of a dense state, then two more layers
sv.submit(layer); sv.submit(layer); sv.flush()
def mid():
    for q in (0, 7, n // 2, n - 1):
        sv.measure_qubit_(q, 0.5)
    sv.submit(layer); sv.flush()
ms, _, st = timed(mid)
out["measure4_then_two_layers_ms"] = round(ms, 2)
out["norm_after"] = sv.norm2()
print(json.dumps(out))
