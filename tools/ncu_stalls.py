#!/usr/bin/env python3
"""Totals of the warp-stall sampling columns of an ncu report (SASS source page) + headline raw metrics."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
hdr = None; tot = collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address": hdr = r; continue
    if hdr and r and r[0].startswith("0x"):
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try: tot[h] += int(r[i])
                except ValueError: pass
s = sum(tot.values())
for k, v in tot.most_common(8): print("%-26s %6.2f%%" % (k, 100 * v / s))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; u = rows[1]
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]:
    if k in h:
        i = h.index(k); print(k, "[%s]" % u[i], [r[i] for r in rows[2:]])
