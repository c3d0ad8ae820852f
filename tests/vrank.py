"""Test helper: P ranks of an in-process rank group (qb_init_group), one Python thread per rank.

With the same device repeated the ranks are VIRTUAL: P shards on one GPU, the real planner
(choose_swaps / swap_schedule), the real fused-pass kernels with rank-bit predicates and the real
pairwise swap kernel (its "peer" pointer aimed at the sibling shard) -- so a 1-GPU box exercises
the P = 2 / 4 / 8 logic.  ctypes releases the GIL inside every C-ABI call, so the ranks really
run concurrently and meet in the library's collectives."""
import threading
import traceback


def run_group(ctxs, fn, timeout=600):
    """fn(rank, ctx) on every rank concurrently; returns the list of results, re-raises the first error."""
    out = [None] * len(ctxs)
    errs = [None] * len(ctxs)

    def work(r):
        try:
            out[r] = fn(r, ctxs[r])
        except BaseException as e:  # noqa: BLE001
            errs[r] = (e, traceback.format_exc())

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(len(ctxs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout)
    hung = [r for r, t in enumerate(ts) if t.is_alive()]
    first = next((e for e in errs if e is not None), None)
    if first is not None:
        raise AssertionError(f"rank failed:\n{first[1]}")
    assert not hung, f"ranks {hung} did not finish"
    return out
