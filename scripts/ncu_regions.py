"""Aggregate an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass)
of k_fused_pass into code regions: samples, executed warp instructions, stall reasons."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
# the file may hold several kernels: split at "Kernel Name" rows; use the first by default
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
print(blk[0][1][:100])
hdr = blk[1]
data = [dict(zip(hdr, r)) for r in blk[2:] if len(r) == len(hdr)]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def classify(ins):
    op = ins.split()[0] if ins.split() else ""
    if op.startswith("@"): op = ins.split()[1]
    return op.split(".")[0]
# regions by instruction type heuristics on a sequential walk
tot = collections.Counter(); inst = collections.Counter(); st = collections.defaultdict(collections.Counter)
for d in data:
    op = classify(d["Source"])
    s = int(d["# Samples"] or 0); e = int(d["Instructions Executed"] or 0)
    tot[op] += s; inst[op] += e
    for c in stall_cols:
        st[op][c] += int(d[c] or 0)
T = sum(tot.values()); E = sum(inst.values())
print(f"total samples {T}, executed warp instructions {E}")
print(f"{'op':10s} {'samp%':>6s} {'exec%':>6s}  top stalls")
for op, s in tot.most_common(22):
    top = ", ".join(f"{k[6:]}={v*100//max(1,s)}%" for k, v in st[op].most_common(4))
    print(f"{op:10s} {100*s/T:6.1f} {100*inst[op]/E:6.1f}  {top}")
allst = collections.Counter()
for op in st:
    allst.update(st[op])
print("all stalls:", ", ".join(f"{k[6:]}={100*v/T:.1f}%" for k, v in allst.most_common(12)))
