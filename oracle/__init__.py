"""CPU oracle for the qubism state-vector hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference (qubitrot/qubism, Haskell +
hmatrix) computes on the one path this repository replaces: gate application,
measurement reduction and collapse on the 2^n ``Complex Double`` amplitude vector.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from here, and only as the checker
(or as the timed CPU baseline) -- never as the product.  ``qubism_b200`` does not
import this package and fails loudly when its CUDA library is missing.

PARITY UNPINNED.  The reference cannot be built in this environment (no GHC /
stack / cabal), its arithmetic lives in the un-vendored ``hmatrix`` dependency
(pinned only through ``resolver: lts-12.4``, stack.yaml:21 -> hmatrix 0.19.x over
system BLAS), and its own test-suite holds no golden amplitudes for this path
(QuickCheck laws on 1-qubit objects only).  The oracle's
authority is therefore a literal reading of the source, cross-validated by
  (i)   ``oracle.dense``  (the reference's real algorithm: dense 2^n x 2^n kron /
        diag / matmul / matvec, exactly as QGate.hs writes it) agreeing with
  (ii)  ``oracle.structured`` (O(2^n) per op, same arithmetic semantics) and with
  (iii) ``oracle/csrc/sv_struct.c`` (the same in C + OpenMP, the CPU baseline),
  (iv)  the reference's QuickCheck properties re-expressed in tests/, and
  (v)   hand-checkable cases (Bell pair of examples/Teleportation.hs:21, ...).

Modules
  dense       literal dense restatement of src/Qubism/QGate.hs + StateVec.hs
  structured  O(2^n) numpy restatement, same semantics, for n up to ~26
  qasm        OpenQASM-2 subset parser + evaluator following
              src/Qubism/QASM/{Parser,Simulation,ProgState}.hs (emits the primitive
              op stream that reaches ``#>`` / ``measureQubit`` / ``collapse``)
  cport       ctypes loader for csrc/sv_struct.c (built by oracle/Makefile)
"""
