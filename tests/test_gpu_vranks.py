"""The sharded path on ONE GPU: P = 2 / 4 / 8 virtual ranks (qb_init_group with the same device
repeated, tests/vrank.py), parity against the oracle through the C ABI.

Everything the multi-GPU run executes is real here -- swap selection (choose_swaps, Belady with the
cyclic tie-break), the exchange schedule, k_fused_pass / qb_jit_pass with rank-bit predicates,
k_peer_swap with the peer pointer aimed at the sibling shard, reductions summed over ranks -- only
the transport differs (host barriers instead of NCCL).  The NCCL transport itself is covered by
tests/test_gpu_dist.py on a multi-GPU box; the host logic alone by tests/test_dist_cpu.py (gloo).
"""
import numpy as np
import pytest

from vrank import run_group

pytestmark = pytest.mark.gpu
TOL = 1e-12  # north_star: amplitudes within 1e-12 absolute of the reference evaluator


@pytest.fixture(scope="module")
def groups():
    import qubism_b200 as Q
    made = {}

    def get(P):
        if P not in made:
            made[P] = Q.Context.group([0] * P)
        return made[P]

    yield get
    for ctxs in made.values():
        for c in ctxs:
            c.close()


def _mixed_ops(n, seed):
    from oracle import dense as D
    from qubism_b200.circuits import qft_ops, random_layers
    ops = random_layers(n, 3, seed=seed, lam0=True)
    ops += [("CU", [0, n - 1], 2, D.unitary(.3, .2, .1)), ("U", 0, np.diag([1, 1j])), ("CX", n - 1, 0), ("CX", 0, 1),
            ("CU", [1], 0, np.diag([1, np.exp(.3j)])), ("CX", 2, 1), ("U", 1, D.unitary(1.1, 2.2, 3.3))]
    return ops + qft_ops(n)


@pytest.mark.parametrize("jit", [0, 1])
@pytest.mark.parametrize("P,n", [(2, 12), (2, 16), (4, 14), (4, 18), (8, 14), (8, 17), (8, 20)])
def test_virtual_ranks_match_the_oracle(groups, P, n, jit):
    """Random layers + controlled / diagonal gates on global qubits + a QFT from a random state;
    then reductions and a measurement on a global qubit, a local one and a swapped one."""
    import qubism_b200 as Q
    from oracle import structured as S
    ctxs = groups(P)
    pbits = P.bit_length() - 1
    L = n - pbits
    rng = np.random.default_rng(1000 + 17 * n + P)
    full = S.gen_state(n, rng)
    ops = _mixed_ops(n, 5 + n)
    ref = S.run_ops(n, ops, full)
    red_ref = {q: S.sumsq(n, q, ref) for q in (0, pbits, n - 1)}
    rb, rv, rp = S.measure_qubit(n, 0, 0.4, ref)

    def rank_fn(r, ctx):
        ctx.set_option("jit", jit)
        ctx.reset_stats()
        sv = Q.StateVec.from_host(full[r << L:(r + 1) << L], n=n, ctx=ctx)
        sv.run_ops(ops)
        got = sv.to_host()
        red = {q: sv.sumsq(q) for q in (0, pbits, n - 1)}
        bit, p = sv.measure_qubit_(0, 0.4)
        got2 = sv.to_host()
        nrm = sv.norm2()
        st = ctx.stats()
        ctx.set_option("jit", 2)
        return got, red, bit, p, got2, nrm, st

    res = run_group(ctxs, rank_fn)
    for r, (got, red, bit, p, got2, nrm, st) in enumerate(res):
        assert np.abs(got - ref).max() < TOL, f"rank {r}"
        for q, (s0, s1) in red.items():
            assert abs(s0 - red_ref[q][0]) < TOL and abs(s1 - red_ref[q][1]) < TOL
        assert bit == rb and abs(p - rp) < TOL
        assert np.abs(got2 - rv).max() < TOL
        assert abs(nrm - 1.0) < TOL
        assert st["exchanges"] >= 1 and st["passes"] >= 1
        if jit == 1 and L >= 10:
            assert st["jit_launches"] >= 1
    # every rank read the same logical amplitudes, bit for bit
    assert all(np.array_equal(res[0][0], x[0]) for x in res[1:])
    if P == 8:  # all three rank bits were needed at once somewhere: a 3-bit all-to-all swap moves 7/8 of a shard
        assert res[0][6]["exchange_bytes"] >= (7 * (16 << L)) // 8


@pytest.mark.parametrize("P", [2, 4, 8])
def test_adder_and_measurements_from_a_fresh_register(groups, P):
    """C4 of BASELINE.json scaled down: the ripple-carry adder (Toffolis via qelib1.inc's ccx) from a
    fresh |0...0> -- ranks other than 0 hold only zeros until a gate reaches a global qubit, so the
    support tracking works across ranks -- then mid-circuit measurements and a random mix of every
    op kind (controlled, diagonal, dense 2- and 3-qubit blocks on global qubits)."""
    import qubism_b200 as Q
    from oracle import structured as S
    from qubism_b200.circuits import adder_ops, random_mixed
    ctxs = groups(P)
    rng = np.random.default_rng(3)
    for k in (6, 8):
        n = 2 * k + 2
        M2 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        M3 = np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0]
        ops = adder_ops(k) + [("MEASURE", 0, 0.3)] + random_mixed(n, 60, 40 + k)
        ops += [("KQ", [0, n - 2], M2 / np.linalg.norm(M2, 2)), ("U", 1, np.array([[0, 1], [1, 0]])), ("KQ", [2, 0, 1], M3),
                ("MEASURE", n - 1, 0.6), ("MEASURE", 1, 0.5)] + random_mixed(n, 30, 90 + k)
        v0 = np.zeros(1 << n, complex)
        v0[0] = 1
        rec_ref = []
        ref = S.run_ops(n, ops, v0, record=rec_ref)

        def rank_fn(r, ctx):
            sv = Q.mkStateVec(n, ctx)
            rec = sv.run_ops(ops)
            return sv.to_host(), rec

        for got, rec in run_group(ctxs, rank_fn):
            assert [(q, b) for q, b, _ in rec] == [(q, b) for q, b, _ in rec_ref]
            assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("P", [2, 8])
def test_repeated_steps_agree_on_the_deferred_factors(groups, P):
    """jit = 2 (the default): structures are compiled in the background at their second sighting, so
    WHICH kernel runs a pass differs from rank to rank and from step to step; the factor a pass is
    defined to leave out must not (exchanges move raw device amplitudes)."""
    import qubism_b200 as Q
    from oracle import structured as S
    from qubism_b200.circuits import random_layers
    ctxs = groups(P)
    n = 18
    L = n - (P.bit_length() - 1)
    ops = random_layers(n, 4, seed=11, lam0=True)
    full = S.gen_state(n, np.random.default_rng(8))
    ref = full
    refs = []
    for _ in range(4):
        ref = S.run_ops(n, ops, ref)
        refs.append(ref)

    def rank_fn(r, ctx):
        sv = Q.StateVec.from_host(full[r << L:(r + 1) << L], n=n, ctx=ctx)
        outs = []
        for step in range(4):
            sv.submit(ops)
            sv.flush()
            if step == 1:
                ctx.jit_wait()
            outs.append(sv.to_host())
        return outs

    for outs in run_group(ctxs, rank_fn):
        for got, want in zip(outs, refs):
            assert np.abs(got - want).max() < TOL


@pytest.mark.parametrize("P", [2, 4])
def test_sharded_vector_space_ops_with_diverged_layouts(groups, P):
    """<.>, +: and tensor between states whose qubit layouts differ (independent global<->local
    swaps), StateVec.hs:51-58,98-100; qb_state_read beyond 2^22 amplitudes."""
    import qubism_b200 as Q
    from oracle import dense as D, structured as S
    ctxs = groups(P)
    pbits = P.bit_length() - 1
    n = 15
    L = n - pbits
    rng = np.random.default_rng(21)
    va, vb = S.gen_state(n, rng), S.gen_state(n, rng)
    opsa = [("U", 0, D.unitary(.3, .2, .1)), ("CX", 0, n - 1), ("U", 1, D.unitary(1., .2, .4))]
    opsb = [("U", n - 1, D.unitary(.5, .1, .0)), ("U", pbits - 1, D.unitary(.7, .9, .2)), ("CX", 3, 0)]
    ra, rb = S.run_ops(n, opsa, va), S.run_ops(n, opsb, vb)
    m = 3  # a small second register, as fuseQRegs tensors in (ProgState.hs:137-166)
    vc = S.gen_state(m + pbits, rng)

    def rank_fn(r, ctx):
        a = Q.StateVec.from_host(va[r << L:(r + 1) << L], n=n, ctx=ctx)
        b = Q.StateVec.from_host(vb[r << L:(r + 1) << L], n=n, ctx=ctx)
        a.run_ops(opsa)
        b.run_ops(opsb)
        dot = a.inner(b)
        dot2 = b.inner(a)
        s = (a + b).to_host()
        Lc = m
        c = Q.StateVec.from_host(vc[r << Lc:(r + 1) << Lc], n=m + pbits, ctx=ctx)
        t = Q.tensor(a, c)
        tv = t.to_host()
        t.apply_1q(0, D.unitary(.1, .2, .3)).apply_cnot(n + m + pbits - 1, 0)
        tv2 = t.to_host()
        return dot, dot2, s, tv, tv2, a.to_host()

    want_t = np.kron(ra, vc)
    nt = n + m + pbits
    want_t2 = S.run_ops(nt, [("U", 0, D.unitary(.1, .2, .3)), ("CX", nt - 1, 0)], want_t)
    for dot, dot2, s, tv, tv2, aa in run_group(ctxs, rank_fn):
        assert abs(dot - np.vdot(ra, rb)) < 1e-12
        assert dot2 == dot.conjugate()  # exact (AlgebraTests.hs:43-47)
        assert np.abs(s - (ra + rb)).max() < TOL
        assert np.abs(aa - ra).max() < TOL  # relayout of an operand does not change what it means
        assert np.abs(tv - want_t).max() < TOL
        assert np.abs(tv2 - want_t2).max() < TOL


def test_read_beyond_two_to_the_22(groups):
    import qubism_b200 as Q
    ctxs = groups(2)
    n = 23
    L = n - 1
    rng = np.random.default_rng(2)
    full = (rng.uniform(-1, 1, 1 << n) + 1j * rng.uniform(-1, 1, 1 << n)) * 2.0 ** -11
    full[5] = -0.0  # a negative zero survives the read

    def rank_fn(r, ctx):
        sv = Q.StateVec.from_host(full[r << L:(r + 1) << L], n=n, ctx=ctx)
        sv.apply_cnot(0, n - 1).apply_cnot(0, n - 1)
        sv.apply_1q(0, np.array([[0, 1], [1, 0]])).apply_1q(0, np.array([[0, 1], [1, 0]]))
        return sv.to_host()

    for got in run_group(ctxs, rank_fn):
        assert np.array_equal(got, full)
        assert np.signbit(got[5].real)


@pytest.mark.parametrize("P,n", [(2, 16), (4, 18), (8, 20)])
def test_swap_carried_by_the_stores_of_a_fused_pass(groups, P, n):
    """Option fuse_exchange (default on): the pass before a global<->local swap writes every tile
    straight into the second shard of the rank that owns it after the swap (XchGeom, qb_internal.h)
    instead of a local store plus an exchange sweep.  Same amplitudes, bit for bit, as the two-step
    path; generic and specialised kernels; the layouts that follow are the same too."""
    import qubism_b200 as Q
    from oracle import structured as S
    from qubism_b200.circuits import random_layers, qft_ops
    ctxs = groups(P)
    L = n - (P.bit_length() - 1)
    ops = random_layers(n, 4, seed=3 + n, lam0=False) + qft_ops(n)
    full = S.gen_state(n, np.random.default_rng(77 + P))
    ref = S.run_ops(n, ops, S.run_ops(n, ops, full))

    def rank_fn(r, ctx):
        outs = {}
        for fuse, jit, tail in ((1, 0, 12), (0, 0, 12), (1, 1, 12), (1, 0, 0)):
            ctx.set_option("fuse_exchange", fuse)
            ctx.set_option("jit", jit)
            ctx.set_option("defer_tail", tail)  # (0: sparse passes at the end of a stuck plan are not held back)
            ctx.reset_stats()
            sv = Q.StateVec.from_host(full[r << L:(r + 1) << L], n=n, ctx=ctx)
            for _ in range(2):
                sv.submit(ops)
                sv.flush()
            outs[(fuse, jit) if tail else "tail0"] = (sv.to_host(), ctx.stats())
        ctx.set_option("fuse_exchange", 1)
        ctx.set_option("jit", 2)
        ctx.set_option("defer_tail", 12)
        return outs

    for outs in run_group(ctxs, rank_fn):
        a, sa = outs[(1, 0)]
        b, sb = outs[(0, 0)]
        c, sc = outs[(1, 1)]
        assert np.abs(a - ref).max() < TOL and np.abs(b - ref).max() < TOL and np.abs(c - ref).max() < TOL
        assert np.array_equal(a, b)
        assert sb["exchanges_fused"] == 0 and sb["exchanges"] >= 2
        assert sa["exchanges"] == sb["exchanges"] and sa["exchange_bytes"] == sb["exchange_bytes"]
        assert sa["exchanges_fused"] >= 1 and sc["exchanges_fused"] == sa["exchanges_fused"]
        assert sa["passes"] == sb["passes"]
        assert sc["jit_launches"] >= 1
        d, sd = outs["tail0"]
        assert np.abs(d - ref).max() < TOL and sd["passes"] >= sa["passes"]
