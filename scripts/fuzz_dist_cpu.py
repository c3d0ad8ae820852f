"""Offline fuzz of the sharded host logic on the CPU (not part of the test suite): swap-carrying
passes (option fuse_exchange) and held-back tail passes (defer_tail) over gloo, world 2 / 4 / 8,
several seeds and thresholds, emulated generic kernel and generated code, against the oracle.
    python scripts/fuzz_dist_cpu.py        # ~5 min; prints one line per case and "bad 0" at the end"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.multiprocessing as mp
import test_dist_cpu as T
if __name__ == "__main__":
    cases = []
    for seed in range(6):
        cases += [(2, 14, 100 + seed, 15, b"tile_bits=10,reg_bits=3", 3, "layers"), (4, 15, 200 + seed, 31, b"tile_bits=10,reg_bits=3,defer_tail=%d" % (seed * 4), 3, "mixed"),
                  (8, 16, 300 + seed, 15, b"tile_bits=10,reg_bits=3,defer_tail=%d" % (3 + seed * 3), 2, "layers")]
    bad = 0
    for world, n, seed, how, opts, nsteps, circ in cases:
        d = tempfile.mkdtemp()
        os.environ["QBE_WORKDIR"] = d
        mp.spawn(T._worker, args=(world, T._free_port(), n, seed, d, how & 15, opts, nsteps, circ), nprocs=world, join=True)
        err, nsw, nbytes, L, nfused, njit = open(os.path.join(d, "result.txt")).read().split()
        ok = float(err) < 1e-12
        bad += not ok
        print(world, n, seed, opts, circ, "err", err, "swaps", nsw, "fused", nfused, "jit", njit, "OK" if ok else "FAIL", flush=True)
    print("bad", bad)
