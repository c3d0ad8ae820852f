"""Exit while specialised kernels are still being compiled in the background: the process must
leave with status 0 (the library drains its compile queue in an atexit handler)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q  # noqa: E402
from qubism_b200.circuits import random_layers  # noqa: E402

n = 20
ops = random_layers(n, 6, seed=9, lam0=True)
sv = Q.mkStateVec(n)
for _ in range(2):  # second sighting: the structures go to the background compiler
    sv.submit(ops)
    sv.flush()
print("leaving with compilations in flight", flush=True)
sys.exit(0)
