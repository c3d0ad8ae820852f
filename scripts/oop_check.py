"""GPU parity of the out-of-place passes (option oop) against the structured oracle, generic and
specialised kernels, plus the layout-aware read / reductions / clones that follow them."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qubism_b200 as Q
from qubism_b200.circuits import qft_ops, random_layers, proper_unitary_layers
from oracle import structured as S

ctx = Q.Context.default()
worst = 0.0
for n in (13, 16, 20):
    for opts in ({"oop": 1, "jit": 0}, {"oop": 1, "jit": 1}, {"oop": 1, "jit": 1, "low_bits": 5}, {"oop": 1, "jit": 0, "lite": 0},
                 {"oop": 1, "jit": 1, "tma": 1}, {"oop": 1, "jit": 1, "tma": 2}):
        for k, v in dict(oop=0, jit=2, low_bits=3, lite=1, tma=0).items():
            ctx.set_option(k, v)
        for k, v in opts.items():
            ctx.set_option(k, v)
        rng = np.random.default_rng(n)
        v = S.gen_state(n, rng)
        ops = qft_ops(n) + random_layers(n, 3, seed=n) + proper_unitary_layers(n, 1)
        ref = S.run_ops(n, ops, v)
        sv = Q.StateVec.from_host(v)
        ctx.reset_stats()
        sv.submit(ops)
        s0, s1 = sv.sumsq(3)
        r0, r1 = S.sumsq(n, 3, ref)
        c = sv.clone()
        c.apply_1q(1, np.array([[0, 1], [1, 0]], complex))
        got = sv.to_host()
        got_c = c.to_host()
        ref_c = S.run_ops(n, [("U", 1, np.array([[0, 1], [1, 0]], complex))], ref)
        st = ctx.stats()
        e = max(float(np.abs(got - ref).max()), float(np.abs(got_c - ref_c).max()), abs(s0 - r0), abs(s1 - r1))
        worst = max(worst, e)
        print(n, opts, "err %.2e" % e, "passes", st["passes"], "jit launches", st["jit_launches"], flush=True)
for k, v in dict(oop=0, jit=2, low_bits=3, lite=1, tma=0).items():
    ctx.set_option(k, v)
assert worst < 1e-12, worst
print("oop parity ok", worst)
