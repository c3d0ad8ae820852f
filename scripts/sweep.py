"""Sweep kernel variants / planner knobs on the benchmark circuit. Usage: sweep.py N [depth]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import random_layers, qft_ops, proper_unitary_layers
n = int(sys.argv[1]); depth = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = Q.Context.default()
sv = Q.mkStateVec(n)
ops = capi.pack_ops(qft_ops(n) + random_layers(n, depth, seed=1000))
opsg = capi.pack_ops(proper_unitary_layers(n, depth))
def run(label, arr, **opts):
    for k, v in opts.items(): ctx.set_option(k, v)
    sv.submit(arr); sv.flush(); ctx.sync(); ctx.reset_stats()
    t0 = time.perf_counter(); sv.submit(arr); sv.flush(); ctx.sync(); dt = time.perf_counter() - t0
    st = ctx.stats()
    print(json.dumps(dict(label=label, opts=opts, ms=round(dt*1e3,2), passes=st["passes"], rounds=st["rounds"], gates=st["ops_executed"],
          ms_per_pass=round(dt*1e3/max(1,st["passes"]),3), gbs=round(st["passes"]*32*(1<<n)/dt/1e9), plan_ms=round(st["plan_ms"],2),
          aups=len(arr)*(1<<n)/dt)), flush=True)
for T, R in ((12,4),(11,4),(10,4),(12,3),(11,3),(10,3),(13,4),(13,5),(12,5)):
    run("qft+rand", ops, tile_bits=T, reg_bits=R, max_rounds=6)
for mr in (3,4,5,8):
    run("qft+rand", ops, tile_bits=12, reg_bits=4, max_rounds=mr)
    run("qft+rand", ops, tile_bits=11, reg_bits=4, max_rounds=mr)
for mg in (8, 12, 16, 24):
    run("qft+rand", ops, tile_bits=12, reg_bits=4, max_rounds=6, max_pass_gates=mg)
ctx.set_option("max_pass_gates", 96)
for T, R in ((12,4),(11,4),(12,3)):
    run("general-class layers", opsg, tile_bits=T, reg_bits=R, max_rounds=6)
