#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed):  python tools/ncu_summary.py rep [kernel-regex]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "local_load_requests", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
last = {}  # the last captured instance of every kernel (warm), in order of first appearance
for r in rows[2:]:
    if len(r) > idx["Kernel Name"]:
        last[r[idx["Kernel Name"]]] = r
for name, r in last.items():
    print("=====", name[:80])
    for k in KEYS:
        if k in idx:
            print("  %-72s %16s %s" % (k, r[idx[k]], units[idx[k]]))
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[idx[h]]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
            except ValueError:
                pass
    print("  stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
