"""A/B of the structure-specialised kernels on the benchmark circuit (QFT-n + 20 random layers):
    python scripts/jit_bench.py 30 "jit=0" "jit=2" "jit=2,lane_fixed=1" ...
Each configuration: 3 warm-up steps (the second one compiles), 3 timed steps (CUDA events)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401  (device init order as in bench.py)
import qubism_b200 as Q  # noqa: E402
from qubism_b200 import capi  # noqa: E402
from qubism_b200.circuits import qft_ops, random_layers  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
configs = sys.argv[2:] or ["jit=0", "jit=2"]
ctx = Q.Context.default()
sv = Q.mkStateVec(n)
ops = capi.pack_ops(qft_ops(n) + random_layers(n, 20, seed=1000))
defaults = {}
for cfg in configs:
    kv = dict(x.split("=") for x in cfg.split(",") if x)
    for k in kv:
        defaults.setdefault(k, ctx.get_option(k))
    for k, v in defaults.items():
        ctx.set_option(k, v)
    for k, v in kv.items():
        ctx.set_option(k, int(v))
    t0 = time.perf_counter()
    for _ in range(3):
        sv.submit(ops)
        sv.flush()
    for _ in range(2):
        ctx.jit_wait()
        sv.submit(ops)
        sv.flush()
    ctx.sync()
    warm = time.perf_counter() - t0
    ctx.reset_stats()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        sv.submit(ops)
        sv.flush()
    ctx.sync()
    ms = (time.perf_counter() - t0) / reps * 1e3
    st = ctx.stats()
    print(f"== {cfg}: {ms:.1f} ms/step  passes {st['passes'] // reps} rounds {st['rounds'] // reps} "
          f"ms/pass {ms / (st['passes'] // reps):.2f}  aups {len(ops) * (1 << n) / ms * 1e3:.3e}  "
          f"jit launches {st['jit_launches']} compiled {st['jit_compiled']} compile_ms {st['jit_compile_ms']:.0f} "
          f"warmup_s {warm:.1f}  norm {sv.norm2():.12f}", flush=True)
