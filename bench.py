#!/usr/bin/env python
"""bench.py -- amplitude-updates/s of the qubism state-vector hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One STEP = one pass of the hot path over one batch of synthetic input: the whole primitive
op stream of the workload circuit applied to an n-qubit device-resident state through the C
ABI (libqubism_sv.so).  One amplitude-update = one amplitude visited by one primitive op of
the reference evaluator's op stream (a U on one qubit or a CX, after qelib1.inc expansion);
a primitive op on an n-qubit state is 2^n amplitude-updates no matter how the backend fuses
or folds it.

Workload (BASELINE.json):  N = 1: 30 qubits, QFT-30 (2,207 ops, examples/fourier.qasm pattern)
followed by 20 random layers (30 U + 15 CX each, 900 ops).  N > 1: the same family on
n = 31 + log2(N) qubits (32 GiB of amplitudes per GPU; N = 8 is the 34-qubit configuration),
sharded one rank per GPU with global<->local qubit swaps over NCCL/NVLink.

The JSON line carries: value (device-timed, inputs resident in HBM), e2e (through the C ABI
from HOST buffers: pinned state upload + one FFI call per gate + result readback inside the
timed region), roofline (HBM, for the fused-pass kernel, measured live with CUDA events on
the launching stream), cpu_baseline (the oracle's C/OpenMP port on the host cores, bounded
sample), clocks (nvidia-smi during the timed region).

--impl reference times the reference's CPU path.  The Haskell reference cannot be built here
(no GHC) and its literal dense algorithm cannot reach 24 qubits on any machine (one gate
matrix would be 4.5 PB), so this arm runs the oracle's structured
C/OpenMP restatement (oracle/csrc/sv_struct.c, kind "port") with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "amplitude_updates_per_s"
UNIT = "amplitude-updates/s"
CPU_SAMPLE_N = 24


def workload(n: int):
    from qubism_b200.circuits import qft_ops, random_layers
    return qft_ops(n) + random_layers(n, 20, seed=1000)


def workload_name(n: int) -> str:
    return f"qft{n}+rand20x{n}: QFT-{n} ({n + 5 * n * (n - 1) // 2 + 2} ops) + 20 random U/CX layers ({20 * (n + n // 2)} ops)"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """The timed region starts now: only samples taken from here on count."""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        for stamp, ln in self.lines:
            if stamp < t0:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_run(n: int, steps: int, warmup: int):
    """Times the oracle's C/OpenMP port (one sweep per primitive op, no fusion) on the host."""
    import numpy as np
    from oracle import cport
    ops = workload(n) if n < CPU_SAMPLE_N else None
    if ops is None:
        from qubism_b200.circuits import qft_ops, random_layers
        ops = qft_ops(n) + random_layers(n, 2, seed=1000)
    packed = cport.pack_ops(ops)
    v = np.zeros(1 << n, dtype=np.complex128)
    v[0] = 1.0
    for _ in range(warmup):
        cport.run_ops_inplace(n, packed, v)
    t0 = time.perf_counter()
    for _ in range(steps):
        cport.run_ops_inplace(n, packed, v)
    dt = time.perf_counter() - t0
    cores = int(cport.lib().sv_num_threads())
    aups = len(ops) * (1 << n) * steps / dt
    sample = (f"QFT-{n} + 2 random layers at n={n} ({len(ops)} primitive ops, {(16 << n) >> 20} MiB state), "
              f"{steps} step(s), one sweep per op, OpenMP x{cores}")
    return aups, dt / steps, cores, sample, len(ops)


def literal_dense_sample():
    """The reference's REAL algorithm (dense kron / matmul, oracle.dense) at the largest size
    that finishes in seconds: shows its O(8^n) wall next to the structured port."""
    import numpy as np
    from oracle import dense as D
    n = 10
    v = D.mkStateVec(n)
    t0 = time.perf_counter()
    nops = 0
    for q in range(3):
        v = D.apply(D.onJust(n, q, D.hadamard()), v)
        v = D.apply(D.cnot(n, q, q + 1), v)
        nops += 2
    dt = time.perf_counter() - t0
    return {"n": n, "ops": nops, "value": nops * (1 << n) / dt, "unit": UNIT,
            "what": "literal dense restatement of QGate.hs:121-154 (numpy/BLAS), 3 H + 3 CX"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    aups, step_s, cores, sample, nops = cpu_run(CPU_SAMPLE_N, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": aups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": workload_name(30 if args.gpus == 1 else 31 + args.gpus.bit_length() - 1),
                   "sample": sample},
        "cpu_baseline": {"value": aups, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": aups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import qubism_b200 as Q
    from qubism_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf.copy_(torch.frombuffer(bytearray(Q.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        ctx = Q.Context(local_rank, rank, world, bytes(idbuf.cpu().numpy().tobytes()))
    else:
        ctx = Q.Context(local_rank)
    n = 30 if world == 1 else 31 + (world.bit_length() - 1)
    n = int(os.environ.get("QB_BENCH_N", n))
    L = n - (world.bit_length() - 1)
    ops = workload(n)
    packed = capi.pack_ops(ops)
    nops = len(ops)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sv = Q.StateVec.create(n, True, ctx)
    ctx.set_option("time_kernels", 1)

    # ---------------- device-resident throughput (`value`)
    def step():
        sv.submit(packed)
        sv.flush()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # nvidia-smi takes a moment to start: launch it before the warm-up
    for _ in range(args.warmup):
        step()
    # pass structures that came back during the warm-up are being compiled into specialised
    # kernels on background threads: let that finish, and give the next sighting (which loads
    # the modules) its own untimed steps -- the timed region measures the steady state
    # (sharded: the layout of the state cycles with a period of a few steps, a structure has to come
    #  round twice before it is compiled)
    for _ in range(2 if world == 1 else 6):
        ctx.jit_wait()
        step()
    barrier()
    ctx.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    elapsed = e0.elapsed_time(e1) / 1e3
    st = ctx.stats()
    if world == 1 and elapsed < 1.5:
        # nvidia-smi samples every 100 ms: keep the same load running (outside the timed region,
        # after the counters were read) until it has seen at least 1.5 s of it
        t_more = time.time()
        while time.time() - t_more < 1.5 - elapsed:
            step()
        ctx.sync()
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = nops * float(1 << n) * args.steps / elapsed
    passes = st["passes"] / args.steps
    fused_ms = st["fused_ms"] / max(1, st["fused_timed"])  # average launch duration
    launches = (st["passes"] + st["simple_launches"] + st["reduce_launches"])

    # ---------------- end to end through the C ABI from host buffers (`e2e`)
    # every step: upload this rank's shard from pinned host memory, one FFI call PER GATE (the way
    # the Haskell interpreter drives the boundary), flush, read back three measurement reductions
    # and a window of amplitudes.
    import ctypes as C
    shard = 1 << L
    chunk = min(shard, 1 << 28)  # 4 GiB of pinned host memory per rank, re-sent to fill the shard
    host = torch.empty(2 * chunk, dtype=torch.float64, pin_memory=True)
    host.zero_()
    one = torch.zeros(2, dtype=torch.float64, pin_memory=True)
    one[0] = 1.0
    win = min(4096, shard)
    ctx.set_option("time_kernels", 0)
    lib, h = ctx.L, sv._h
    calls = []  # the per-gate FFI calls, arguments marshalled once (a compiled host pays ~100 ns each)
    for op in ops:
        if op[0] == "U":
            calls.append((lib.qb_apply_1q, (h, op[1], capi.mat4(op[2]))))
        else:
            calls.append((lib.qb_apply_cnot, (h, op[1], op[2])))

    def e2e_step():
        for off in range(0, shard, chunk):
            sv.write_local((host.data_ptr(), chunk), first=off)
        if rank == 0:
            sv.write_local((one.data_ptr(), 1), first=0)
        for fn, a in calls:
            rc = fn(*a)
            if rc != 0:
                capi.check(rc)
        sv.flush()
        red = [sv.sumsq(q) for q in (0, n // 2, n - 1)]
        w = sv.local_to_host(0, win)
        return red, w

    e2e_step()
    ctx.jit_wait()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        red, w = e2e_step()
    barrier()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = nops * float(1 << n) / e2e_dt
    h2d = 16 * shard + nops * 80
    d2h = 3 * 16 + 16 * win

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (k_fused_pass), HBM-bound
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    alg_bytes = 32.0 * float(1 << L)  # 16 B read + 16 B written per local amplitude per pass
    achieved = alg_bytes / (fused_ms / 1e3) / 1e9 if fused_ms > 0 else None
    traffic = None
    for tag in ("r01e", "r01d"):  # newest ncu --set full capture of the dominant kernel
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_fused_pass_summary.json")))
            traffic = prof.get("dram_bytes_per_launch_scaled_to", {}).get(str(L))
            break
        except (OSError, ValueError):
            continue
    jit_share = st["jit_launches"] / max(1, st["passes"])
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": ("qb_jit_pass (k_fused_pass specialised per pass structure with NVRTC)" if jit_share > 0.5
                           else "k_fused_pass"),
                "specialised_share_of_launches": jit_share,
                "launches_per_step": passes, "avg_launch_ms": fused_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": (st["fused_ms"] / 1e3) / elapsed if elapsed > 0 else None,
                # what the same op stream would move one sweep per primitive op: may
                # exceed the HBM peak, which is the point of fusing
                "unfused_equivalent_gbs": value / world * 32.0 / 1e9}
    if world > 1:
        # T_roof(P) = passes x 32 B x 2^L / HBM  +  bytes sent per GPU / NVLink per direction
        nvl = 770.0  # GB/s per direction, measured peer copy on this pool (900 nominal)
        xb = st["exchange_bytes"] / args.steps
        t_roof = passes * alg_bytes / (peak * 1e9) + xb / (nvl * 1e9)
        roofline["sharded"] = {"t_roof_ms": t_roof * 1e3, "t_measured_ms": elapsed / args.steps * 1e3,
                               "efficiency": t_roof / (elapsed / args.steps),
                               "hbm_term_ms": passes * alg_bytes / (peak * 1e9) * 1e3, "nvlink_term_ms": xb / (nvl * 1e9) * 1e3,
                               "nvlink_gbs_per_direction": nvl,
                               "what": "SURVEY.md 8d: sum of the passes at the HBM peak plus the exchanged bytes at the "
                                       "measured NVLink rate, over the measured step"}

    # ---------------- CPU baseline (rank 0, N = 1 only): the oracle's C port, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        aups, _, cores, sample, _ = cpu_run(CPU_SAMPLE_N, 1, 0)
        cpu = {"value": aups, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "literal_dense": literal_dense_sample()}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": workload_name(n), "qubits": n, "local_qubits": L, "primitive_ops_per_step": nops,
                   "state_bytes_per_gpu": 16 << L, "l2_policy": "inputs larger than L2 (state >= 16 GiB >> 126 MB)",
                   "parallelism": f"shard{world}" if world > 1 else "single",
                   "planner": {k: ctx.get_option(k) for k in ("tile_bits", "reg_bits", "low_bits", "lane_fixed", "max_rounds", "peephole", "rot", "lite", "jit")},
                   "specialised_kernels": {"compiled": st["jit_compiled"], "compile_ms_total": st["jit_compile_ms"],
                                           "launches_in_timed_region": st["jit_launches"],
                                           "note": "pass structures seen twice are compiled with NVRTC on background threads during "
                                                   "the warm-up steps (W + 2 untimed steps); the timed steps hit the cache"},
                   "ops_executed_per_step": st["ops_executed"] / args.steps,
                   "ops_folded_per_step": st["ops_folded"] / args.steps,
                   "passes_per_step": passes, "rounds_per_step": st["rounds"] / args.steps,
                   "exchanges_per_step": st["exchanges"] / args.steps,
                   "exchange_bytes_per_gpu_per_step": st["exchange_bytes"] / args.steps,
                   "plan_ms_per_step": st["plan_ms"] / args.steps},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_dt * 1e3, "ffi_calls_per_step": nops + 6,
                "check": {"s1_q0": red[0][1], "amp0": [float(w[0].real), float(w[0].imag)]}},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least three warm-up steps
        run_ours(args)


if __name__ == "__main__":
    main()
