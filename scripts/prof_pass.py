"""Profiling target: a few fused passes of a chosen shape.  Usage:
   python scripts/prof_pass.py N TILE REG WORKLOAD [REPS]
WORKLOAD: low12g (12 general gates on the 12 low bits), low12r (real class), one (1 gate),
          rand (2 random layers)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200.circuits import random_layers
from qubism_b200.qgate import unitary_matrix

n, T, R, wl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ctx = Q.Context.default()
ctx.set_option("tile_bits", T); ctx.set_option("reg_bits", R)
if wl in ("real52", "gen13"): ctx.set_option("peephole", 0)
for kv in os.environ.get("QB_OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); ctx.set_option(k, int(v))
sv = Q.mkStateVec(n)
G = unitary_matrix(0.3, 0.2, 0.1); U = unitary_matrix(0.3, 0.2, 0.0)
def build():
    if wl == "low12g":
        for q in range(n - 12, n): sv.apply_1q(q, G)
    elif wl == "low12r":
        for q in range(n - 12, n): sv.apply_1q(q, U)
    elif wl == "one":
        sv.apply_1q(0, G)
    elif wl == "one0":
        sv.apply_1q(n - 1, G)
    elif wl == "real52":
        for _ in range(13):
            for q in range(4): sv.apply_1q(q, U)
    elif wl == "gen13":
        for _ in range(13): sv.apply_1q(0, G)
    elif wl == "rand":
        sv.submit(random_layers(n, 2, seed=1000))
build(); sv.flush(); ctx.sync()
ctx.reset_stats()
t0 = time.perf_counter()
for _ in range(reps):
    build(); sv.flush()
ctx.sync()
dt = (time.perf_counter() - t0) / reps
st = ctx.stats()
p = st["passes"] / reps
print(f"n={n} T={T} R={R} {wl}: {dt*1e3:.3f} ms/rep, passes/rep={p}, rounds/rep={st['rounds']/reps}, "
      f"{p*32*(1<<n)/dt/1e9:.0f} GB/s per pass")
