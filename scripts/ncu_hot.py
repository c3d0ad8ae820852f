"""Summarise an ncu report's SASS source page: samples / executed counts by opcode and the hottest lines.
   usage: ncu_hot.py report.ncu-rep [kernel_index]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; blocks.append(cur); continue
    if cur is not None: cur.append(r)
b = blocks[kidx]; hdr = b[0]; data = b[1:]
iA, iE, iS = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
byop, samp, ins, tot, stot = Counter(), Counter(), [], 0, 0
for r in data:
    try: e, s = int(r[iE]), int(r[iS])
    except (ValueError, IndexError): continue
    t = r[iA].strip(); p = t.split()
    op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
    byop[op] += e; samp[op] += s; tot += e; stot += s; ins.append((s, e, t))
print(f"kernel {kidx}: {tot:.3e} warp instructions, {stot} samples")
for op, c in byop.most_common(24):
    print(f"  {op:10s} exec {100*c/tot:5.1f}%   samples {100*samp[op]/stot:5.1f}%")
ins.sort(reverse=True)
print("hottest lines:")
for s, e, t in ins[:30]:
    print(f"  {100*s/stot:5.2f}%  exec={e:11d}  {t[:100]}")
