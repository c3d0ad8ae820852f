"""First-contact GPU script: small parity checks + raw pass timings at 30 qubits.
Run under gpurun; writes gpurun_out/smoke.json."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200.circuits import random_layers, qft_ops
from oracle import structured as S, dense as D

out = {}
ctx = Q.Context.default()
rng = np.random.default_rng(7)

def parity(n, ops, opts=None, tag=""):
    for k, v in (opts or {}).items():
        ctx.set_option(k, v)
    v0 = S.gen_state(n, rng)
    sv = Q.StateVec.from_host(v0)
    sv.run_ops(ops)
    got = sv.to_host()
    ref = S.run_ops(n, ops, v0)
    d = float(np.abs(got - ref).max())
    print(f"parity n={n} {tag} opts={opts} maxdiff={d:.3e}", flush=True)
    return d

res = []
for n in (1, 2, 5, 9):
    ops = random_layers(n, 2, seed=3, lam0=False) if n > 1 else [("U", 0, D.unitary(1, 2, 3))]
    res.append(("simple", n, parity(n, ops)))
for n in (10, 12, 14, 16):
    ops = random_layers(n, 3, seed=n, lam0=True) + [("CU", [0, n - 1], 3, D.unitary(.3, .2, .1)), ("U", n - 1, np.diag([1, 1j])),
                                                    ("CU", [2], n - 2, np.diag([1, np.exp(.3j)])), ("U", 0, D.unitary(0, 0, .7))]
    for opts in ({"tile_bits": 12, "reg_bits": 4}, {"tile_bits": 12, "reg_bits": 5}, {"tile_bits": 11, "reg_bits": 4},
                 {"tile_bits": 13, "reg_bits": 5}, {"tile_bits": 13, "reg_bits": 4}, {"tile_bits": 12, "reg_bits": 3},
                 {"tile_bits": 10, "reg_bits": 3}, {"tile_bits": 11, "reg_bits": 5}, {"tile_bits": 11, "reg_bits": 3}, {"tile_bits": 10, "reg_bits": 4}):
        res.append(("fused", n, opts, parity(n, ops, opts)))
ctx.set_option("tile_bits", 12); ctx.set_option("reg_bits", 4)
# measurement
n = 14
v0 = S.gen_state(n, rng)
sv = Q.StateVec.from_host(v0)
for q in (0, 5, 13):
    s = sv.sumsq(q); r = S.sumsq(n, q, v0)
    print("sumsq", q, s, r, flush=True)
    res.append(("sumsq", q, abs(s[0] - r[0]), abs(s[1] - r[1])))
bit, p = sv.measure_qubit_(3, 0.4)
b2, v1, p2 = S.measure_qubit(n, 3, 0.4, v0)
res.append(("measure", bit, b2, abs(p - p2), float(np.abs(sv.to_host() - v1).max())))
print(res[-1], flush=True)
out["parity"] = res

# ---- raw pass timings at n = 30 (16 GiB) -------------------------------------------------
n = int(os.environ.get("SMOKE_N", "30"))
sv = Q.mkStateVec(n)
H = D.hadamard(); U = D.unitary(0.3, 0.2, 0.0); G = D.unitary(0.3, 0.2, 0.1)
timings = []
def timed(label, build, reps=3, **opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    build(); sv.flush(); ctx.sync()
    ctx.reset_stats()
    t0 = time.perf_counter()
    for _ in range(reps):
        build(); sv.flush()
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    st = ctx.stats()
    passes = st["passes"] / reps
    gbs = passes * 32 * (1 << n) / dt / 1e9
    rec = dict(label=label, opts=opts, ms=dt * 1e3, passes=passes, rounds=st["rounds"] / reps, gbs_per_pass=gbs,
               gates=st["ops_executed"] / reps)
    print(json.dumps(rec), flush=True)
    timings.append(rec)

for T, R in ((12, 4), (12, 5), (11, 4), (13, 5), (13, 4), (12, 3), (11, 5), (11, 3), (10, 3), (10, 4)):
    timed("1 general gate, high bit (1 round)", lambda: sv.apply_1q(0, G), tile_bits=T, reg_bits=R)
for T, R in ((12, 4), (12, 5), (11, 4), (13, 5)):
    timed("1 general gate, bit 0 (3 rounds)", lambda: sv.apply_1q(n - 1, G), tile_bits=T, reg_bits=R)
    def layer():
        for q in range(n - 12, n): sv.apply_1q(q, G)
    timed("12 general gates on the 12 low bits", layer, tile_bits=T, reg_bits=R)
    def layer_r():
        for q in range(n - 12, n): sv.apply_1q(q, U)
    timed("12 real-class gates on the 12 low bits", layer_r, tile_bits=T, reg_bits=R)
    def hi():
        for q in range(0, 7): sv.apply_1q(q, G)
    timed("7 general gates on the 7 high bits", hi, tile_bits=T, reg_bits=R)
for lb in (3, 5, 6, 7):
    def hi():
        for q in range(0, 4): sv.apply_1q(q, G)
    timed("4 general gates on high bits, low_bits=%d" % lb, hi, tile_bits=12, reg_bits=4, low_bits=lb)
ctx.set_option("low_bits", 5); ctx.set_option("tile_bits", 12); ctx.set_option("reg_bits", 4)
def circ():
    sv.submit(qft_ops(n) + random_layers(n, 20, seed=1000))
timed("QFT + 20 random layers", circ, reps=1)
t0 = time.perf_counter(); s = sv.sumsq(3); dt = time.perf_counter() - t0
timings.append(dict(label="sumsq", ms=dt * 1e3, gbs=16 * (1 << n) / dt / 1e9)); print(timings[-1], flush=True)
out["timings"] = timings
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/smoke.json", "w"), indent=1, default=str)
print("SMOKE DONE")
