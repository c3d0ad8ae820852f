"""Per-pass CUDA-event times of the benchmark step beside each pass's tile positions (in-place plan,
specialised kernels compiled at first sight).  python scripts/pass_times.py [n] [options]"""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
log = "/tmp/qb_pass_log.txt"
os.environ["QB_PASS_LOG"] = log
import torch  # noqa
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import qft_ops, random_layers
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
opts = sys.argv[2] if len(sys.argv) > 2 else ""
ctx = Q.Context.default()
for kv in [x for x in opts.split(",") if x]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
ctx.set_option("jit", 1)
ctx.set_option("time_kernels", 1)
oplist = qft_ops(n) + random_layers(n, 20, seed=1000)
ops = capi.pack_ops(oplist)
sv = Q.mkStateVec(n)
sv.submit(ops); sv.flush(); ctx.sync(); ctx.stats()
open(log, "w").close()
sv.submit(ops); sv.flush(); ctx.sync(); ctx.stats()
ms = [float(x) for x in open(log).read().split()]
txt = capi.plan_describe(n, oplist, opts)
rows = re.findall(r"pass (\d+) .*?tile=\[([0-9,]+)\] rounds=(\d+) gates=(\d+).*?types\[general,real,diag,swap,rot,general1\]=([0-9,]+)", txt)
print("passes", len(ms), "planned", len(rows), "total ms", sum(ms))
for i, t in enumerate(ms):
    if i < len(rows):
        pos = [int(x) for x in rows[i][1].split(",")]
        pages = sum(1 for p in pos if p >= 17)
        print(f"pass {i:2d} {t:7.3f} ms  rounds {rows[i][2]} gates {rows[i][3]:>2s} types {rows[i][4]:14s} page-bits {pages} tile {rows[i][1]}")
    else:
        print(f"pass {i:2d} {t:7.3f} ms")
