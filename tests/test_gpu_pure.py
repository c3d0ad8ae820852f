"""Value semantics behind the boundary: the reference's pure `g #> sv` (QGate.hs:78-80), used by
the interpreter once per primitive op (`sv' = g idx #> sv; writeStateVec sv'`,
QASM/Simulation.hs:94-122), must keep fusion.  qb_state_clone is lazy: a clone is a handle on
the same shard at the same log position; ops between two observations fuse however they were
sliced into handles; data is copied only when an OLDER live value shares the shard, and that
copy rides on the first fused pass."""
import gc

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _circuit(n):
    from qubism_b200.circuits import qft_ops, random_layers
    return qft_ops(n) + random_layers(n, 6, seed=1000)


def _inplace_stats(ctx, n, ops, v):
    import qubism_b200 as Q
    ctx.reset_stats()
    sv = Q.StateVec.from_host(v, ctx=ctx)
    sv.run_ops(ops)
    out = sv.to_host()
    return out, ctx.stats()


@pytest.mark.parametrize("free_old", [True, False], ids=["old_value_finalised_at_once", "old_values_alive_until_gc"])
def test_with_index_pattern_keeps_fusion(ctx, default_opts, free_old):
    """QFT-20 + layers through the withIndex pattern: clone -> apply ONE op -> (drop the old value).
    With the old values finalised at once nothing is ever copied; if the garbage collector is late
    (every old value still alive at the observation) there is exactly one copy-on-write, fused
    into the first pass.  Either way the pass count equals the in-place plan's."""
    import qubism_b200 as Q
    from oracle import structured as S
    n = 20
    ops = _circuit(n)
    v = S.gen_state(n, np.random.default_rng(4))
    ref = S.run_ops(n, ops, v)
    ctx.set_option("jit", 0)
    base, st0 = _inplace_stats(ctx, n, ops, v)
    assert np.abs(base - ref).max() < TOL
    ctx.reset_stats()
    sv = Q.StateVec.from_host(v, ctx=ctx)
    first = sv
    keep = []
    for op in ops:
        new = sv.apply_pure([op])
        if free_old and sv is not first:
            sv.free()
        else:
            keep.append(sv)
        sv = new
    got = sv.to_host()
    st = ctx.stats()
    assert np.abs(got - ref).max() < TOL
    assert np.array_equal(got, base), "same plan, same kernels: bitwise equal to the in-place run"
    assert st["clones"] == len(ops)
    assert st["passes"] == st0["passes"] <= 2 * st0["passes"]
    assert st["cow_fused"] == 1 and st["cow_copies"] == 0  # `first` (the uploaded value) is still alive
    # the original value is untouched, and so is every intermediate one that is still alive
    assert np.abs(first.to_host() - v).max() == 0.0
    if not free_old:
        k = len(ops) // 2
        mid = keep[k].to_host()  # = ops[:k] applied to v
        assert np.abs(mid - S.run_ops(n, ops[:k], v)).max() < TOL
        assert np.abs(sv.to_host() - ref).max() < TOL
    del keep
    gc.collect()


def test_linear_option_consumes_old_values(ctx, default_opts):
    """Option "linear": older handles of the lineage are consumed by an observation of a newer one --
    no copy at all, and a later use of a consumed value fails loudly instead of reading wrong data."""
    import qubism_b200 as Q
    from oracle import structured as S
    from qubism_b200.capi import QbError, QB_ERR_STATE
    n = 16
    ops = _circuit(n)
    v = S.gen_state(n, np.random.default_rng(5))
    ctx.set_option("linear", 1)
    try:
        ctx.reset_stats()
        sv0 = Q.StateVec.from_host(v, ctx=ctx)
        sv = sv0
        olds = []
        for op in ops:
            olds.append(sv)
            sv = sv.apply_pure([op])
        got = sv.to_host()
        st = ctx.stats()
        assert np.abs(got - S.run_ops(n, ops, v)).max() < TOL
        assert st["cow_fused"] == 0 and st["cow_copies"] == 0
        with pytest.raises(QbError) as ei:
            olds[3].to_host()
        assert ei.value.code == QB_ERR_STATE
    finally:
        ctx.set_option("linear", 0)


def test_branching_values_stay_independent(ctx):
    """measureQubit keeps qr, collapse One qr and collapse Zero qr alive at once
    (StateVec.hs:121-129): two different continuations of one value."""
    import qubism_b200 as Q
    from oracle import dense as D, structured as S
    n = 14
    v = S.gen_state(n, np.random.default_rng(6))
    qr = Q.StateVec.from_host(v, ctx=ctx)
    qr.apply_1q(3, D.unitary(.4, .5, .6))           # queued on the shared lineage
    base = S.apply_1q(n, 3, D.unitary(.4, .5, .6), v)
    one = Q.collapse(2, 1, qr)
    zero = Q.collapse(2, 0, qr)
    a = qr.clone().apply_1q(0, D.hadamard()).apply_cnot(0, 5)
    b = qr.clone().apply_1q(0, D.pauliX())
    assert np.abs(one.to_host() - S.collapse(n, 2, 1, base)).max() < TOL
    assert np.abs(zero.to_host() - S.collapse(n, 2, 0, base)).max() < TOL
    assert np.abs(a.to_host() - S.apply_cnot(n, 0, 5, S.apply_1q(n, 0, D.hadamard(), base))).max() < TOL
    assert np.abs(b.to_host() - S.apply_1q(n, 0, D.pauliX(), base)).max() < TOL
    assert np.abs(qr.to_host() - base).max() < TOL
    # vector-space results are new values too; their operands survive
    s = a + b
    assert np.abs(s.to_host() - (a.to_host() + b.to_host())).max() < TOL
    y = a.clone()
    y.axpy_(2.0, y)  # x and y share a shard
    assert np.abs(y.to_host() - 3.0 * a.to_host()).max() < 4 * TOL


def test_measurement_through_pure_clones(ctx):
    """The interpreter's observe: runStateT (measureQubit k) sv on the value in the map, one qubit
    at a time (Simulation.hs:136-144) -- every step clones, measures, and drops the old value late."""
    import qubism_b200 as Q
    from oracle import structured as S
    n = 13
    v = S.gen_state(n, np.random.default_rng(7))
    rs = np.random.default_rng(8).uniform(0, 1, n)
    sv = Q.StateVec.from_host(v, ctx=ctx)
    ref = v
    keep = []
    for q in range(n):
        new = sv.clone()
        bit, p = new.measure_qubit_(q, rs[q])
        rb, ref, rp = S.measure_qubit(n, q, rs[q], ref)
        assert bit == rb and abs(p - rp) < TOL
        keep.append(sv)
        sv = new
    assert np.abs(sv.to_host() - ref).max() < TOL
    assert np.abs(keep[0].to_host() - v).max() == 0.0


@pytest.mark.parametrize("n", [13, 16])
def test_dense_block_then_gates_in_one_flush_from_a_known_support(ctx, n):
    """A dense block populates its qubits: the fused segment that follows it IN THE SAME FLUSH must
    not plan its dead tiles from the support the block has just invalidated (round-1 advisor
    finding: zeros(n) -> apply_kq -> 1q gates lost the gates on the block's new amplitudes)."""
    import qubism_b200 as Q
    from oracle import dense as D, structured as S
    rng = np.random.default_rng(n)
    M = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    ops = [("KQ", [0, 1], M), ("U", 5, D.hadamard()), ("CX", 0, n - 1), ("U", n - 2, D.unitary(.3, .4, .5)),
           ("KQ", [n - 1, 3, 6], np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]),
           ("U", 6, D.hadamard()), ("CX", 3, 2)]
    v0 = np.zeros(1 << n, complex)
    v0[0] = 1
    sv = Q.mkStateVec(n, ctx)
    sv.run_ops(ops)
    assert np.abs(sv.to_host() - S.run_ops(n, ops, v0)).max() < TOL
    # ... and the same after a collapse (one known bit)
    v = S.gen_state(n, rng)
    pre = [("COLLAPSE", 1, 1), ("COLLAPSE", 4, 0)]
    sv = Q.StateVec.from_host(v, ctx=ctx)
    sv.run_ops(pre)
    sv.to_host()
    sv.run_ops(ops)
    assert np.abs(sv.to_host() - S.run_ops(n, pre + ops, v)).max() < TOL


def test_free_after_shutdown_and_stale_contexts():
    """A finalizer may run after qb_shutdown (ForeignPtr finalizers run at GC time)."""
    import qubism_b200 as Q
    from qubism_b200.capi import QbError
    c = Q.Context(0)
    sv = Q.mkStateVec(10, c)
    sv.apply_1q(0, np.eye(2))
    c.close()
    with pytest.raises(QbError):
        sv.to_host()
    sv.free()  # no crash
