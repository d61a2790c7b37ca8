#!/usr/bin/env python
"""Per-SOURCE-LINE instruction counts and stall samples of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on):  python tools/ncu_lines.py rep launch-index [n]"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, out = None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        print(r[1])
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > ie and r[0].isdigit():
        try:
            out.append((int(r[ie] or 0), int(r[sm] or 0), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot, smp = sum(o[0] for o in out), sum(o[1] for o in out)
print("warp instructions %d, samples %d" % (tot, smp))
for ie, sm_, f, ln, src in sorted(out, reverse=True)[:n]:
    print("%5.1f%% instr  %5.1f%% smp  %s:%d  %s" % (100.0 * ie / tot, 100.0 * sm_ / max(smp, 1), f, ln, src))
