#!/usr/bin/env python
"""Top stall locations of one kernel in an .ncu-rep:  python tools/ncu_hot.py rep launch-index [n] [stall column]

launch-index = position of the kernel in the report (ncu -i rep --page raw --csv lists them)."""
import csv, io, subprocess, sys
from collections import Counter
rep, skip = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][:2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {}
for i, h in enumerate(hdr):
    idx.setdefault(h, i)
data = [r for r in rows[hi + 1:] if len(r) > idx["Instructions Executed"] and r[idx["# Samples"]].isdigit()]
seen, uniq = set(), []
for r in data:
    if r[0] in seen:
        continue
    seen.add(r[0]); uniq.append(r)
data = uniq
tot = sum(int(r[idx["# Samples"]]) for r in data)
ins = sum(int(r[idx["Instructions Executed"]]) for r in data)
print("samples", tot, "warp-instructions", ins, "sass lines", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "(" not in h]
tots = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
print("stall totals:", ", ".join("%s %d" % (c[6:], v) for c, v in sorted(tots.items(), key=lambda kv: -kv[1]) if v))
byop, iop = Counter(), Counter()
for r in data:
    op = [o for o in r[idx["Source"]].split() if not o.startswith("@")][0].split(".")[0]
    byop[op] += int(r[idx["# Samples"]]); iop[op] += int(r[idx["Instructions Executed"]])
for op, c in byop.most_common(16):
    print("  %-8s samples %6d (%4.1f%%)  instr %10d (%4.1f%%)" % (op, c, 100 * c / tot, iop[op], 100 * iop[op] / ins))
print()
for i, r in enumerate(data):
    r.append(i)
# 64-instruction windows: samples, executed, dominant stall
print("window  samples  exec/1k  top stalls")
for w0 in range(0, len(data), 64):
    ws = data[w0:w0 + 64]
    s = sum(int(r[idx["# Samples"]]) for r in ws)
    e = sum(int(r[idx["Instructions Executed"]]) for r in ws)
    if not e:
        continue
    st = {c: sum(int(r[idx[c]] or 0) for r in ws) for c in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print("#%5d  %6d  %8d  smp/kexe %.2f  %s" % (w0, s, e // 1000, 1000.0 * s / e, " ".join("%s:%d" % (c[6:], v) for c, v in top if v)))
print()
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%5s smp %8s exe  #%5d  %-70s %s" % (r[idx["# Samples"]], r[idx["Instructions Executed"]], r[-1], r[idx["Source"]].strip()[:70], st))
if len(sys.argv) > 4:
    col = sys.argv[4]
    print("\n== %s ==" % col)
    for r in sorted(data, key=lambda r: -int(r[idx[col]] or 0))[:20]:
        print("%5s  #%5d  %s" % (r[idx[col]], r[-1], r[idx["Source"]].strip()[:80]))
