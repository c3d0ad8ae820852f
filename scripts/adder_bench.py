"""C4: examples/rippleCarryAdder.qasm widened to 15-bit operands on one 32-qubit
register (Toffolis = qelib1.inc ccx: 9 U + 6 CX), sharded over the GPUs of the job.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/adder_bench.py [k]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch
import torch.distributed as dist
import qubism_b200 as Q
from qubism_b200 import capi
from qubism_b200.circuits import adder_ops, random_layers

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idbuf.copy_(torch.frombuffer(bytearray(Q.Context.unique_id()), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    ctx = Q.Context(lr, rank, world, bytes(idbuf.cpu().numpy().tobytes()))
else:
    ctx = Q.Context(lr)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 15
n = 2 * k + 2
ops = adder_ops(k)
packed = capi.pack_ops(ops)
out = {"n": n, "world": world, "primitive_ops": len(ops)}
def timed(f):
    ctx.barrier(); ctx.reset_stats(); t0 = time.perf_counter(); f(); ctx.barrier()
    return (time.perf_counter() - t0) * 1e3, ctx.stats()
# (a) from a fresh |0...0>: the answer is a basis state, the support keeps most tiles dead
sv = Q.mkStateVec(n, ctx)
ms, st = timed(lambda: (sv.submit(packed), sv.flush()))
s0, s1 = sv.sumsq(2 * k + 1)  # cout
out["fresh_ms"] = round(ms, 2); out["fresh_tiles"] = st["tiles"]; out["fresh_passes"] = st["passes"]; out["cout_s1"] = s1
# (b) on a dense state (one random layer first): every tile is live
sv.submit(capi.pack_ops(random_layers(n, 1, seed=3))); sv.flush()
ms, st = timed(lambda: (sv.submit(packed), sv.flush()))
out["dense_ms"] = round(ms, 2); out["dense_passes"] = st["passes"]; out["dense_exchanges"] = st["exchanges"]
out["dense_aups"] = len(ops) * float(1 << n) / (ms / 1e3)
out["ops_executed"] = st["ops_executed"]; out["ops_folded"] = st["ops_folded"]
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    ctx.barrier(); dist.destroy_process_group()
