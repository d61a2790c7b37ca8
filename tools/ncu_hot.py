#!/usr/bin/env python
"""Top stall locations of one kernel in an .ncu-rep:  python tools/ncu_hot.py rep kernel-regex [n]"""
import csv, io, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) > idx["Instructions Executed"] and r[idx["# Samples"]].isdigit()]
# the csv repeats every row twice (two views); keep unique addresses
seen, uniq = set(), []
for r in data:
    if r[0] in seen:
        continue
    seen.add(r[0]); uniq.append(r)
data = uniq
tot = sum(int(r[idx["# Samples"]]) for r in data)
ins = sum(int(r[idx["Instructions Executed"]]) for r in data)
print("samples", tot, "warp-instructions", ins, "sass lines", len(data))
byop, iop = Counter(), Counter()
for r in data:
    op = [o for o in r[idx["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    byop[op] += int(r[idx["# Samples"]]); iop[op] += int(r[idx["Instructions Executed"]])
for op, c in byop.most_common(14):
    print("  %-8s samples %6d (%4.1f%%)  instr %10d (%4.1f%%)" % (op, c, 100 * c / tot, iop[op], 100 * iop[op] / ins))
print()
for i, r in enumerate(data):
    r.append(i)
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    print("%5s smp %8s exe  #%5d  %s" % (r[idx["# Samples"]], r[idx["Instructions Executed"]], r[-1], r[idx["Source"]][:90]))

# per-stall-reason hot spots (optional 4th arg: stall column, e.g. stall_no_inst)
if len(sys.argv) > 4:
    col = sys.argv[4]
    tot_c = sum(int(r[idx[col]] or 0) for r in data)
    print("\n== %s: %d samples ==" % (col, tot_c))
    # cluster by 64-instruction windows to see where in the code they sit
    win = Counter()
    for r in data:
        win[r[-1] // 64] += int(r[idx[col]] or 0)
    for w_, c in sorted(win.items()):
        if c:
            print("  sass #%5d-%5d : %5d  %s" % (w_ * 64, w_ * 64 + 63, c, "#" * (c * 200 // max(tot_c, 1))))
    for r in sorted(data, key=lambda r: -int(r[idx[col]] or 0))[:15]:
        print("%5s  #%5d  %s" % (r[idx[col]], r[-1], r[idx["Source"]][:80]))
