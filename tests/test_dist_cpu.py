"""Multi-rank host logic on the CPU: world_size 2, 4 and 8 over the gloo backend.

Each rank owns a shard (numpy), runs the SAME planner / swap-selection / exchange-schedule
code as the NCCL path (choose_swaps, swap_schedule, apply_swaps_to_perm in qb_planner.cpp)
through the test-only emulator, and moves half-shards with torch.distributed send/recv.  The
reassembled state must equal the oracle's.  The CUDA kernels themselves are covered by the
-m gpu tests; the real NCCL exchange by `gpurun --gpus N`."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, out_dir, any_local, options=b"", nsteps=1, circuit="layers"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import structured as S
    from qubism_b200 import capi
    from qubism_b200.circuits import random_layers, random_mixed
    from oracle import dense as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E = C.CDLL(os.path.join(ROOT, "tests", "emul", "libqb_emul.so"))
    XCHG = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64)
    E.qbe_run_rank.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(capi.QbOp), C.c_int64, C.c_char_p, C.c_void_p,
                               C.POINTER(C.c_int), XCHG, C.POINTER(C.c_int64), C.c_int]
    nbytes = [0]

    def xchg(peer, send, recv, nd):
        s = torch.from_numpy(np.ctypeslib.as_array(send, shape=(nd,)).copy())
        r = torch.empty(nd, dtype=torch.float64)
        reqs = [dist.isend(s, peer), dist.irecv(r, peer)]
        for q in reqs:
            q.wait()
        np.ctypeslib.as_array(recv, shape=(nd,))[:] = r.numpy()
        nbytes[0] += nd * 8
        return 0

    pbits = world.bit_length() - 1
    L = n - pbits
    rng = np.random.default_rng(seed)
    full = S.gen_state(n, rng)
    ops = random_layers(n, 3, seed=seed, lam0=True) + [("CU", [0, n - 1], 3, D.unitary(.3, .2, .1)),
                                                      ("U", 0, np.diag([1, 1j])), ("CX", n - 1, 0), ("CX", 0, 1),
                                                      ("CU", [1], 0, np.diag([1, np.exp(.3j)]))]
    if circuit == "mixed":  # every op kind the ABI takes, multi-controlled and general (non-unitary) matrices included
        ops = random_mixed(n, 90, seed)
    shard = np.ascontiguousarray(full[rank << L:(rank + 1) << L]).copy()
    perm = (C.c_int * n)(*range(n))
    arr = capi.pack_ops(ops)
    st = (C.c_int64 * 6)()
    cb = XCHG(xchg)
    nsw = nfused = njit = 0
    for _ in range(nsteps):  # (an iterated circuit: the later steps start from the layout the earlier ones left)
        rc = E.qbe_run_rank(n, world, rank, arr, len(arr), options, shard.ctypes.data_as(C.c_void_p), perm, cb, st, any_local)
        assert rc >= 1, f"expected at least one global<->local swap, rc={rc}"
        nsw += rc
        nfused += st[4]
        njit += st[5]
    # every rank must have made the same layout decisions
    perms = [None] * world
    dist.all_gather_object(perms, list(perm))
    assert all(p == perms[0] for p in perms)
    np.save(os.path.join(out_dir, f"shard{rank}.npy"), shard)
    dist.barrier()
    if rank == 0:
        phys = np.concatenate([np.load(os.path.join(out_dir, f"shard{r}.npy")) for r in range(world)])
        idx = np.arange(1 << n)
        pidx = np.zeros_like(idx)
        for q in range(n):
            pidx |= ((idx >> q) & 1) << perms[0][q]
        got = phys[pidx]
        ref = full
        for _ in range(nsteps):
            ref = S.run_ops(n, ops, ref)
        err = float(np.abs(got - ref).max()) / max(1.0, float(np.abs(ref).max()))
        with open(os.path.join(out_dir, "result.txt"), "w") as f:
            f.write(f"{err} {nsw} {nbytes[0]} {L} {nfused} {njit}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,any_local", [(2, 13, 0), (2, 13, 1), (2, 13, 3), (4, 13, 0), (4, 13, 1), (4, 13, 3),
                                               (8, 14, 0), (8, 14, 3)],
                         ids=lambda v: str(v))
def test_sharded_exchange_over_gloo(tmp_path, emul, world, n, any_local):
    """any_local = 0: the top local bits are evicted (contiguous blocks, what the NCCL send/recv
    fallback moves); 1: the bits needed furthest in the future (what the peer-memory kernel moves)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 77 + world, str(tmp_path), any_local), nprocs=world, join=True)
    err, nsw, nbytes, L, _, _ = open(tmp_path / "result.txt").read().split()
    assert float(err) < 1e-13
    # volume per swap of k bits is (1 - 2^-k) of the shard each way
    assert int(nbytes) <= int(nsw) * 16 * (1 << int(L))


@pytest.mark.parametrize("world,n,options,nsteps,how", [(2, 13, b"tile_bits=10,reg_bits=3", 2, 7), (4, 14, b"tile_bits=10,reg_bits=3", 2, 15),
                                                        (8, 16, b"tile_bits=10,reg_bits=3", 2, 7), (8, 15, b"", 3, 15),
                                                        (2, 18, b"", 4, 15), (4, 15, b"defer_tail=0", 3, 7),
                                                        (8, 16, b"defer_tail=20,tile_bits=10,reg_bits=3", 3, 7),
                                                        (4, 15, b"", 3, 23), (8, 16, b"tile_bits=10,reg_bits=3", 2, 31)],
                         ids=lambda v: str(v))
def test_swap_carried_by_the_stores_of_the_last_pass(tmp_path, emul, world, n, options, nsteps, how):
    """Option fuse_exchange: the out-of-place pass before a global<->local swap stores every tile
    straight into the shard of the rank that owns it afterwards.  The emulator runs the planner's
    geometry (fused_exchange_geometry -> XchGeom) and the kernels' destination arithmetic per tile,
    with one buffer per destination rank and gloo in NVLink's place; nothing lands outside the
    places a rank owns in its peers' shards, and several steps in a row (each starts from the
    layout the one before left) reassemble to the oracle's state.  how = 7: the emulated generic
    kernel; 15: the GENERATED code of that pass (host flavour of the specialised kernel's source,
    its peer table aimed at the per-rank buffers); + 16: a random mix of every op kind instead of
    rotation / CX layers."""
    os.environ["QBE_WORKDIR"] = str(tmp_path)
    circuit = "mixed" if how & 16 else "layers"
    how &= 15
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 31 + world, str(tmp_path), how, options, nsteps, circuit), nprocs=world,
             join=True)
    err, nsw, nbytes, L, nfused, njit = open(tmp_path / "result.txt").read().split()
    assert float(err) < 1e-13
    assert int(nfused) >= 1, "no swap was carried by a pass"
    # (passes that hold controlled-U gates are not specialised: the mixed circuits may never reach the generated code)
    assert how != 15 or circuit == "mixed" or int(njit) >= 1, "the generated code never ran"
    assert int(nbytes) <= int(nsw) * 16 * (1 << int(L))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_repeated_steps_settle_into_a_layout_cycle(world):
    """An iterated circuit on a sharded state.  Out-of-place passes re-sort the local qubits by next use
    after every pass (ties by qubit label), the swaps' tie-break looks at the same op stream again
    (qb_planner.cpp choose_swaps, `future`): the logical->physical layout -- local order AND the set of
    qubits on the rank bits -- comes back after a few steps at 2, 4 and 8 ranks, so the
    structure-specialised kernels find their pass structures compiled.  (In place the 8-rank layout never
    came back: three rank bits keep permuting the local positions.)  Planner only: no amplitudes."""
    import ctypes as C
    import subprocess
    from qubism_b200 import capi
    from qubism_b200.circuits import qft_ops, random_layers
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d, "libqb_emul.so"], stdout=subprocess.DEVNULL)
    E = C.CDLL(os.path.join(d, "libqb_emul.so"))
    E.qbe_layout_trace.argtypes = [C.c_int, C.c_int, C.POINTER(capi.QbOp), C.c_int64, C.c_char_p, C.c_int, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    n = 31 + world.bit_length() - 1
    nsteps = 20
    ops = capi.pack_ops(qft_ops(n) + random_layers(n, 20, seed=1000))
    perm = (C.c_int * (nsteps * n))()
    cnt = (C.c_int64 * (nsteps * 3))()
    assert E.qbe_layout_trace(n, world, ops, len(ops), b"", nsteps, 3, perm, cnt) == 0  # 3 = Belady + cyclic lookahead
    layouts = [tuple(perm[s * n:(s + 1) * n]) for s in range(nsteps)]
    period = next(p for p in range(1, 7) if all(layouts[s] == layouts[s - p] for s in range(13, nsteps)))
    assert period <= 6  # (1 at 2 ranks, 6 at 4, 3 at 8 on this circuit)
    assert all(cnt[3 * s + 2] == 0 for s in range(12, nsteps)), "new pass structures keep appearing"
    # option defer_tail (default 12): sparse passes at the end of a stuck plan wait for the plan after the
    # swap -- fewer sweeps per step than with every schedulable gate run as early as possible
    steady = sum(cnt[3 * s] for s in range(14, nsteps))
    swaps = sum(cnt[3 * s + 1] for s in range(14, nsteps))
    assert E.qbe_layout_trace(n, world, ops, len(ops), b"defer_tail=0", nsteps, 3, perm, cnt) == 0
    assert steady < sum(cnt[3 * s] for s in range(14, nsteps))
    assert swaps <= sum(cnt[3 * s + 1] for s in range(14, nsteps)) + (nsteps - 14)  # (at most one more swap per step)
    # the in-place schedule (oop = 0) still cycles at 2 and 4 ranks
    if world <= 4:
        assert E.qbe_layout_trace(n, world, ops, len(ops), b"oop=0", 10, 3, perm, cnt) == 0
        assert all(cnt[3 * s + 2] == 0 for s in range(7, 10))
