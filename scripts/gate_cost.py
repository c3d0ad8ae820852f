"""Per-gate marginal cost by class (slope between 1 and K gates in one pass) + circuit totals."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import random_layers, qft_ops, proper_unitary_layers
from qubism_b200.qgate import unitary_matrix
n = int(sys.argv[1]); K = 13
ctx = Q.Context.default()
sv = Q.mkStateVec(n)
G = unitary_matrix(0.3, 0.2, 0.1); U = np.array([[0.8, -0.6], [0.6, 0.8]]); Dg = np.diag([np.exp(.1j), np.exp(.7j)]); X = np.array([[0, 1], [1, 0]])
def t_pass(build, reps=3):
    build(); sv.flush(); ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(reps): build(); sv.flush()
    ctx.sync(); dt = (time.perf_counter() - t0) / reps
    st = ctx.stats()
    return dt * 1e3, st["passes"] / reps, st["rounds"] / reps
def slope(label, one):
    ctx.set_option("peephole", 0)
    t1, p1, r1 = t_pass(lambda: one(1))
    tk, pk, rk = t_pass(lambda: one(K))
    ctx.set_option("peephole", 1)
    print(json.dumps(dict(label=label, ms_1=round(t1, 3), ms_K=round(tk, 3), passes=(p1, pk), rounds=(r1, rk), per_gate_ms=round((tk - t1) / (K - 1), 4))), flush=True)
for T, R in ((12, 4), (11, 4)):
    ctx.set_option("tile_bits", T); ctx.set_option("reg_bits", R)
    print("== T,R", T, R, flush=True)
    slope("general hi", lambda k: [sv.apply_1q(0, G) for _ in range(k)])
    slope("real hi", lambda k: [sv.apply_1q(0, U) for _ in range(k)])
    slope("diag hi", lambda k: [sv.apply_1q(0, Dg) for _ in range(k)])
    slope("cx t=hi c=reg", lambda k: [sv.apply_cnot(1, 0) for _ in range(k)])
    slope("cx t=hi c=thr(low)", lambda k: [sv.apply_cnot(n - 1, 0) for _ in range(k)])
    slope("cx t=hi c=ext", lambda k: [sv.apply_cnot(10, 0) for _ in range(k)])
    slope("general bit0 (3 rounds)", lambda k: [sv.apply_1q(n - 1, G) for _ in range(k)])
    slope("real 4 distinct hi bits xK", lambda k: [sv.apply_1q(q, U) for _ in range(k) for q in range(4)])
    slope("general 12 low bits xK", lambda k: [sv.apply_1q(q, G) for _ in range(k) for q in range(n - 12, n)][:80])
    slope("real 12 low bits xK", lambda k: [sv.apply_1q(q, U) for _ in range(k) for q in range(n - 12, n)][:80])
ops = capi.pack_ops(qft_ops(n) + random_layers(n, 20, seed=1000))
opsg = capi.pack_ops(proper_unitary_layers(n, 20))
def run(label, arr, **opts):
    for k, v in opts.items(): ctx.set_option(k, v)
    sv.submit(arr); sv.flush(); ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter(); sv.submit(arr); sv.flush(); ctx.sync(); dt = time.perf_counter() - t0
    st = ctx.stats()
    print(json.dumps(dict(label=label, opts=opts, ms=round(dt*1e3,2), passes=st["passes"], rounds=st["rounds"],
          ms_per_pass=round(dt*1e3/max(1,st["passes"]),3), gbs=round(st["passes"]*32*(1<<n)/dt/1e9), aups=len(arr)*(1<<n)/dt)), flush=True)
for T, R in ((12,4),(11,4),(10,4),(12,3)):
    run("qft+rand", ops, tile_bits=T, reg_bits=R)
run("general", opsg, tile_bits=12, reg_bits=4)
run("general", opsg, tile_bits=11, reg_bits=4)
