"""The specialised fused pass with bulk-asynchronous tile loads (option "tma": cp.async.bulk into the
transpose buffer one tile ahead, mbarrier completion, round 0 reads its registers from shared
memory) against the oracle and against the register-load flavour of the same kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("n,opts", [(13, {}), (16, {}), (16, {"reg_bits": 5}), (16, {"tile_bits": 11, "reg_bits": 3}),
                                    (18, {"max_pass_gates": 6}), (20, {}), (20, {"low_bits": 5}), (20, {"l2_prefetch": 0})])
def test_bulk_async_loads_match_the_oracle(ctx, default_opts, n, opts):
    import qubism_b200 as Q
    from oracle import structured as S
    from qubism_b200.circuits import qft_ops, random_layers
    ops = qft_ops(n) + random_layers(n, 5, seed=70 + n, lam0=True)
    v = S.gen_state(n, np.random.default_rng(n))
    ref = S.run_ops(n, ops, v)
    for k, val in opts.items():
        ctx.set_option(k, val)
    ctx.set_option("jit", 1)
    outs = {}
    for tma in (0, 1, 2):
        ctx.set_option("tma", tma)
        ctx.reset_stats()
        sv = Q.StateVec.from_host(v, ctx=ctx)
        sv.submit(ops)
        outs[tma] = sv.to_host()
        st = ctx.stats()
        assert st["jit_launches"] == st["passes"] > 0
        assert np.abs(outs[tma] - ref).max() < TOL
    # the same arithmetic in the same order: only the way the tile reaches the registers differs
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("tma", [1, 2])
def test_bulk_async_loads_through_lazy_clones_and_from_known_support(ctx, default_opts, tma):
    """Out-of-place first pass (copy-on-write) and dead-tile skipping with the bulk-copy pipeline."""
    import qubism_b200 as Q
    from oracle import structured as S
    from qubism_b200.circuits import random_layers
    n = 18
    ops = random_layers(n, 4, seed=9, lam0=True)
    ctx.set_option("jit", 1)
    ctx.set_option("tma", tma)
    v = S.gen_state(n, np.random.default_rng(3))
    sv = Q.StateVec.from_host(v, ctx=ctx)
    ctx.reset_stats()
    new = sv.apply_pure(ops)
    got = new.to_host()
    assert ctx.stats()["cow_fused"] == 1
    assert np.abs(got - S.run_ops(n, ops, v)).max() < TOL
    assert np.array_equal(sv.to_host(), v)
    z = np.zeros(1 << n, complex)
    z[0] = 1
    sv = Q.mkStateVec(n, ctx)
    sv.submit(ops)
    assert np.abs(sv.to_host() - S.run_ops(n, ops, z)).max() < TOL


@pytest.mark.parametrize("tma", [1, 2])
def test_round_trip_26_qubits_with_bulk_async_loads(ctx, default_opts, tma):
    """Size-independent property at a size the oracle cannot reach: C^-1 C |0> = |0>."""
    import qubism_b200 as Q
    from qubism_b200.circuits import random_layers
    n = 26
    ops = random_layers(n, 6, seed=5, lam0=True)
    inv = []
    for op in reversed(ops):
        inv.append(op if op[0] == "CX" else ("U", op[1], np.asarray(op[2]).conj().T))
    ctx.set_option("tma", tma)
    ctx.set_option("jit", 1)
    sv = Q.mkStateVec(n, ctx)
    sv.apply_1q(0, np.array([[1, 1], [1, -1]]) / np.sqrt(2))  # leave the fully known support first
    for q in range(1, n):
        sv.apply_cnot(q - 1, q)
    sv.flush()
    sv.submit(ops)
    sv.flush()
    sv.submit(inv)
    for q in reversed(range(1, n)):
        sv.apply_cnot(q - 1, q)
    sv.apply_1q(0, np.array([[1, 1], [1, -1]]) / np.sqrt(2))
    head = sv.to_host(0, 1 << 16)
    assert abs(head[0] - 1.0) < TOL and np.abs(head[1:]).max() < TOL
    assert abs(sv.norm2() - 1.0) < TOL
