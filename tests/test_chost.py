"""The drop-in boundary from a host that is neither Python nor C++: tests/chost/teleport.c, plain
C99 over include/qubism_sv.h, linked against libqubism_sv.so the way a `foreign import ccall`
binding links.  Here: the header is valid C99 (-pedantic -Werror), every symbol the
program uses resolves, and without a device the program ends with the library's loud
"no CPU fallback" error instead of computing anything.  On the GPU: it teleports a qubit
(examples/Teleportation.hs) through the interpreter's per-op pure `#>` pattern and checks a product
state of `unitary theta phi lambda` gates against its closed form."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "chost", "teleport.c")


def _build(tmp_path):
    from qubism_b200 import capi
    capi.lib()  # (raises if the CUDA library has not been built)
    exe = str(tmp_path / "teleport")
    libdir = os.path.join(ROOT, "qubism_b200")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC,
           "-o", exe, "-L", libdir, "-lqubism_sv", "-lm", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return exe


def test_c_host_builds_links_and_fails_loudly_without_a_device(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a device is present: the run is test_c_host_on_the_gpu")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stderr
    assert "qb_init" in r.stderr and "ok" not in r.stdout


@pytest.mark.gpu
def test_c_host_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("ok")
    assert " 0 separate copies" in r.stdout  # the per-op pure pattern never copied the state
