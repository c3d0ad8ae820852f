"""Out-of-place passes (option `oop`, the single-GPU default): every tile is written as one
contiguous block of a second shard and the whole qubit layout is re-sorted by next use after each
pass.  The layout is an implementation detail behind the C ABI: amplitudes come back in INDEX order
(StateVec.hs:60-68), reductions, collapse, clones and the vector-space operations between states
whose layouts diverged (StateVec.hs:51-58,98-100) all see the reference's qubit numbering.  Checked
against the structured oracle; the planner side (layouts depend on the op stream only, an iterated
circuit finds its structures again) is covered on the CPU in tests/test_host_cpu.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _circuit(n, layers=4, seed=7):
    from qubism_b200.circuits import qft_ops, random_layers, proper_unitary_layers
    return qft_ops(n) + random_layers(n, layers, seed=seed) + proper_unitary_layers(n, 1)


@pytest.mark.parametrize("knobs", [dict(oop=1), dict(oop=2), dict(oop=1, jit=0), dict(oop=1, lite=0, jit=0), dict(oop=1, oop_low_bits=3),
                                   dict(oop=1, oop_low_bits=7), dict(oop=1, chunk_lanes=1), dict(oop=1, tma=1), dict(oop=1, tma=2),
                                   dict(oop=1, reg_bits=5), dict(oop=1, tile_bits=11), dict(oop=1, tile_bits=13), dict(oop=0)],
                         ids=lambda k: ",".join(f"{a}={b}" for a, b in k.items()))
@pytest.mark.parametrize("n", [12, 15, 19])
def test_parity_with_every_knob(ctx, default_opts, n, knobs):
    import qubism_b200 as Q
    from oracle import structured as S
    if knobs.get("tile_bits", 12) > n:
        pytest.skip("tile larger than the state")
    ctx.set_option("jit", 1)
    for k, v in knobs.items():
        ctx.set_option(k, v)
    ops = _circuit(n)
    v = S.gen_state(n, np.random.default_rng(n))
    ref = S.run_ops(n, ops, v)
    ctx.reset_stats()
    sv = Q.StateVec.from_host(v, ctx=ctx)
    sv.submit(ops)
    got = sv.to_host()
    st = ctx.stats()
    assert np.abs(got - ref).max() < TOL
    assert st["passes"] >= 2 and st["simple_launches"] <= 1, "the fused kernels ran (one extra launch at most: the index-order gather)"
    # a range read in the middle of the index space, and the reductions, in the reference's numbering
    lo, cnt = (1 << n) // 3, 1000
    assert np.array_equal(sv.to_host(lo, cnt), got[lo:lo + cnt])
    for q in (0, n // 2, n - 1):
        s0, s1 = sv.sumsq(q)
        r0, r1 = S.sumsq(n, q, ref)
        assert abs(s0 - r0) < TOL and abs(s1 - r1) < TOL


def test_measurement_and_more_gates_on_a_moved_layout(ctx, default_opts):
    """collapse / measureQubit (StateVec.hs:104-129) after the layout has moved, then more gates: the
    dead-tile passes that follow a collapse run in place on the moved layout."""
    import qubism_b200 as Q
    from oracle import structured as S
    n = 16
    ctx.set_option("jit", 1)
    ops = _circuit(n, 3)
    v = S.gen_state(n, np.random.default_rng(3))
    sv = Q.StateVec.from_host(v, ctx=ctx)
    sv.submit(ops)
    ref = S.run_ops(n, ops, v)
    for q, r in ((2, 0.3), (n - 1, 0.9), (7, 0.01)):
        bit, pone = sv.measure_qubit_(q, r)
        rbit, ref, rp = S.measure_qubit(n, q, r, ref)  # (bit, collapsed state, pOne)
        assert bit == rbit and abs(pone - rp) < TOL
        more = _circuit(n, 1, seed=100 + q)
        sv.submit(more)
        ref = S.run_ops(n, more, ref)
    assert np.abs(sv.to_host() - ref).max() < TOL


def test_algebra_between_states_with_different_layouts(ctx, default_opts):
    """<.>, +: and tensor (StateVec.hs:51-58,98-100) between states that different circuits left in
    different layouts; clones share a shard until written (copy-on-write rides on an out-of-place pass)."""
    import qubism_b200 as Q
    from oracle import structured as S
    n = 14
    ctx.set_option("jit", 1)
    rng = np.random.default_rng(5)
    va, vb = S.gen_state(n, rng), S.gen_state(n, rng)
    oa, ob = _circuit(n, 3, seed=1), _circuit(n, 2, seed=2)[::-1]
    a = Q.StateVec.from_host(va, ctx=ctx)
    b = Q.StateVec.from_host(vb, ctx=ctx)
    a.submit(oa)
    b.submit(ob)
    ra, rb = S.run_ops(n, oa, va), S.run_ops(n, ob, vb)
    assert abs(a.inner(b) - np.vdot(ra, rb)) < 1e-12
    c = a.clone()
    c.submit(ob)  # the older value `a` stays alive: copy-on-write
    assert np.abs(c.to_host() - S.run_ops(n, ob, ra)).max() < TOL
    assert np.abs(a.to_host() - ra).max() < TOL
    s = a + b
    assert np.abs(s.to_host() - (ra + rb)).max() < TOL
    small = Q.StateVec.from_host(S.gen_state(3, rng), ctx=ctx)
    t = Q.tensor(a, small)
    assert np.abs(t.to_host() - np.kron(ra, small.to_host())).max() < TOL
    # a partial upload into a state whose layout has moved: the untouched amplitudes keep their meaning
    patch = S.gen_state(4, rng)
    a.write_local(patch, first=32)
    want = ra.copy()
    want[32:48] = patch
    assert np.abs(a.to_host() - want).max() < TOL


def test_an_iterated_circuit_finds_its_specialised_kernels_again(ctx, default_opts):
    """The layouts out-of-place passes produce depend on the op stream only: from the third step on,
    a repeated circuit plans the same structures (period <= 2) and every pass of it runs a
    specialised kernel (policy jit = 2: compiled at the second sighting)."""
    import qubism_b200 as Q
    from oracle import structured as S
    n = 18
    ctx.set_option("jit", 2)
    ops = _circuit(n, 5)
    v = S.gen_state(n, np.random.default_rng(9))
    sv = Q.StateVec.from_host(v, ctx=ctx)
    ref = v
    for step in range(8):
        ctx.jit_wait()
        ctx.reset_stats()
        sv.submit(ops)
        sv.flush()
        st = ctx.stats()
        ref = S.run_ops(n, ops, ref)
    assert st["jit_launches"] == st["passes"], f"{st['jit_launches']} of {st['passes']} passes specialised in step 8"
    got = sv.to_host()
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, float(np.abs(ref).max()))
