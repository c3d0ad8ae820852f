"""Debug matrix for option fuse_exchange on virtual ranks: per step max |err| against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import qubism_b200 as Q  # noqa: E402
from oracle import structured as S  # noqa: E402
from qubism_b200.circuits import random_layers  # noqa: E402
from vrank import run_group  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
ctxs = Q.Context.group([0] * P)
L = n - (P.bit_length() - 1)
ops = random_layers(n, 4, seed=11, lam0=True)
full = S.gen_state(n, np.random.default_rng(8))
refs = []
ref = full
for _ in range(4):
    ref = S.run_ops(n, ops, ref)
    refs.append(ref)
for fuse in (0, 1):
    for jit in (0, 1, 2):
        def rank_fn(r, ctx):
            ctx.set_option("fuse_exchange", fuse)
            ctx.set_option("jit", jit)
            ctx.reset_stats()
            sv = Q.StateVec.from_host(full[r << L:(r + 1) << L], n=n, ctx=ctx)
            errs = []
            for step in range(4):
                sv.submit(ops)
                sv.flush()
                if step == 1:
                    ctx.jit_wait()
                errs.append(float(np.abs(sv.to_host() - refs[step]).max()))
            st = ctx.stats()
            return errs, {k: st[k] for k in ("passes", "exchanges", "exchanges_fused", "jit_launches", "jit_compiled")}
        for r, (errs, st) in enumerate(run_group(ctxs, rank_fn)):
            print(f"P={P} n={n} fuse={fuse} jit={jit} rank={r} errs={['%.1e' % e for e in errs]} {st}", flush=True)
