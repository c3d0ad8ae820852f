"""Kernel time of the benchmark step with specialised kernels compiled at first sight (jit = 1):
CUDA-event time of the fused passes only, so the NVRTC time of never-recurring structures does not count.
   python scripts/oop_time.py 30 "oop=0" "oop=1,low_bits=5" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import qft_ops, random_layers
n = int(sys.argv[1])
ctx = Q.Context.default()
sv = Q.mkStateVec(n)
ops = capi.pack_ops(qft_ops(n) + random_layers(n, 20, seed=1000))
defaults = {}
for cfg in sys.argv[2:]:
    kv = dict(x.split("=") for x in cfg.split(",") if x)
    for k in kv: defaults.setdefault(k, ctx.get_option(k))
    for k, v in defaults.items(): ctx.set_option(k, v)
    for k, v in kv.items(): ctx.set_option(k, int(v))
    ctx.set_option("jit", 1)
    ctx.set_option("time_kernels", 1)
    sv.submit(ops); sv.flush(); ctx.sync()
    for rep in range(2):
        ctx.reset_stats()
        sv.submit(ops); sv.flush(); ctx.sync()
        st = ctx.stats()
        print(f"== {cfg}: fused {st['fused_ms']:.1f} ms in {st['passes']} passes = {st['fused_ms'] / st['passes']:.2f} ms/pass, "
              f"jit launches {st['jit_launches']}, norm {sv.norm2():.12f}", flush=True)
    ctx.set_option("time_kernels", 0)
