"""Multi-GPU parity check: run under torchrun, one rank per GPU.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_check.py [n]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import qubism_b200 as Q
from oracle import dense as D, structured as S
from qubism_b200.circuits import random_layers, qft_ops

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idbuf.copy_(torch.frombuffer(bytearray(Q.Context.unique_id()), dtype=torch.uint8))
dist.broadcast(idbuf, 0)
ctx = Q.Context(lr, rank, world, bytes(idbuf.cpu().numpy().tobytes()))
pbits = world.bit_length() - 1
ok = True
for n in ([int(sys.argv[1])] if len(sys.argv) > 1 else [pbits + 4, 14, 18]):
    L = n - pbits
    rng = np.random.default_rng(99)
    full = S.gen_state(n, rng)
    shard = full[rank << L:(rank + 1) << L]
    sv = Q.StateVec.from_host(shard, n=n, ctx=ctx)
    ops = random_layers(n, 3, seed=5, lam0=True) + [("CU", [0, n - 1], 2, D.unitary(.3, .2, .1)), ("U", 0, np.diag([1, 1j])),
                                                   ("CX", n - 1, 0), ("CX", 0, 1), ("CU", [1], 0, np.diag([1, np.exp(.3j)]))]
    ops += qft_ops(n)
    ctx.reset_stats()
    sv.run_ops(ops)
    got = sv.to_host(0, min(1 << n, 1 << 20))
    ref = S.run_ops(n, ops, full)
    err = float(np.abs(got - ref[:got.size]).max())
    st = ctx.stats()
    # reductions and measurement on a global qubit, a local one, and after the layout has changed
    red_err = 0.0
    for q in (0, pbits, n - 1):
        s = sv.sumsq(q); r = S.sumsq(n, q, ref)
        red_err = max(red_err, abs(s[0] - r[0]), abs(s[1] - r[1]))
    bit, p = sv.measure_qubit_(0, 0.4)
    rb, rv, rp = S.measure_qubit(n, 0, 0.4, ref)
    got2 = sv.to_host(0, min(1 << n, 1 << 20))
    merr = float(np.abs(got2 - rv[:got2.size]).max())
    nrm = sv.norm2()
    good = err < 1e-12 and red_err < 1e-12 and bit == rb and abs(p - rp) < 1e-12 and merr < 1e-12 and abs(nrm - 1) < 1e-12
    ok = ok and good
    if rank == 0:
        print(f"n={n} world={world}: amp err {err:.2e}, sumsq err {red_err:.2e}, measure bit {bit}=={rb} err {merr:.2e}, "
              f"exchanges {st['exchanges']} ({st['exchange_bytes']/2**20:.1f} MiB/rank), passes {st['passes']}, simple {st['simple_launches']} -> "
              f"{'OK' if good else 'FAIL'}", flush=True)
# C4: the widened ripple-carry adder (Toffolis through qelib1.inc's ccx), started
# from a fresh |0...0> -- the support is fully known, ranks other than 0 hold only zeros until a
# gate reaches a global qubit -- followed by a random mix of every op kind and two measurements
from qubism_b200.circuits import adder_ops, random_mixed
for k in (6, 8):
    n = 2 * k + 2
    if n - pbits < 4:
        continue
    ops = adder_ops(k) + [("MEASURE", 0, 0.3)] + random_mixed(n, 60, 40 + k) + [("MEASURE", n - 1, 0.6), ("MEASURE", 1, 0.5)]
    v0 = np.zeros(1 << n, complex)
    v0[0] = 1
    rec_ref = []
    ref = S.run_ops(n, ops, v0, record=rec_ref)
    sv = Q.mkStateVec(n, ctx)
    ctx.reset_stats()
    rec = sv.run_ops(ops)
    got = sv.to_host(0, 1 << n)
    err = float(np.abs(got - ref).max())
    bits_ok = [(q, b) for q, b, _ in rec] == [(q, b) for q, b, _ in rec_ref]
    st = ctx.stats()
    good = err < 1e-11 and bits_ok
    ok = ok and good
    if rank == 0:
        print(f"adder k={k} (n={n}) + mixed from |0>: amp err {err:.2e}, measured bits match {bits_ok}, exchanges {st['exchanges']}, "
              f"passes {st['passes']}, tiles {st['tiles']} -> {'OK' if good else 'FAIL'}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
ctx.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK", "PASSED" if int(flag.item()) else "FAILED", flush=True)
sys.exit(0 if int(flag.item()) else 1)
