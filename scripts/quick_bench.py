import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import random_layers, qft_ops, proper_unitary_layers
from qubism_b200.qgate import unitary_matrix
n=int(sys.argv[1]) if len(sys.argv)>1 else 30
ctx=Q.Context.default(); sv=Q.mkStateVec(n)
G=unitary_matrix(.3,.2,.1)
def t(build, reps=3):
    build(); sv.flush(); ctx.sync(); ctx.reset_stats()
    t0=time.perf_counter()
    for _ in range(reps): build(); sv.flush()
    ctx.sync(); return (time.perf_counter()-t0)/reps*1e3
print("1 gate pass ms", round(t(lambda: sv.apply_1q(0,G)),3), " bit0 (3 rounds) ms", round(t(lambda: sv.apply_1q(n-1,G)),3))
ops=capi.pack_ops(qft_ops(n)+random_layers(n,20,seed=1000))
ms=t(lambda: sv.submit(ops), reps=2)
st=ctx.stats()
print("circuit ms", round(ms,1), "passes", st["passes"]//2, "rounds", st["rounds"]//2, "ms/pass", round(ms/(st["passes"]//2),2), "aups %.3e" % (len(ops)*(1<<n)/ms*1e3))
opsg=capi.pack_ops(proper_unitary_layers(n,20))
ms=t(lambda: sv.submit(opsg), reps=1)
print("general-class circuit ms", round(ms,1))
