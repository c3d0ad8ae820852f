"""Sharded path on real GPUs: one process per GPU over NCCL (torchrun), parity against the
oracle through qb_state_read.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a 1-GPU box.
The host-side exchange logic is covered on the CPU by tests/test_dist_cpu.py (gloo)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("jit", [2, 1])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_parity_over_nccl(world, jit):
    """jit = 1: every step pass runs as a structure-specialised kernel on every rank; the factors
    those kernels leave out must agree across shards (exchanges move raw device amplitudes)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "scripts", "dist_check.py")], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, QB_JIT=str(jit)))
    assert out.returncode == 0 and "DIST CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
