#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by out-of-line device function (the UR3E_PHASE functions of the step kernel).

usage: ncu_by_func.py <report.ncu-rep> <cubin> <ncu-kernel-name-substring> <cubin-section-substring>
Joins the per-instruction counters of `ncu --page source --print-source sass --csv` with the function labels of
`nvdisasm -c` on the same cubin (instruction order is identical).
"""
import csv, io, re, subprocess, sys, collections
rep, cubin, kern, sect = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
kernels, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name": cur = dict(name=r[1], hdr=None, rows=[]); kernels.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address": cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r and r[0].startswith("0x"): cur["rows"].append(r)
k = [x for x in kernels if kern in x["name"]][0]
hdr = k["hdr"]; ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
infn, fn, seq = False, "?", []
for ln in dis.split("\n"):
    if ln.startswith("//--------------------- .text."): infn = sect in ln; fn = "kernel"; continue
    if not infn: continue
    m = re.match(r"^(\$?[_A-Za-z$][\w$]*):\s*$", ln)
    if m and not m.group(1).startswith(".L"):
        name = m.group(1)
        mms = list(re.finditer(r"\$_ZN4ur3e(\d+)", name))
        if mms:
            mm = mms[-1]; i = mm.end(); fn = name[i:i + int(mm.group(1))]
        else: fn = "kernel"
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): seq.append(fn)
n = min(len(seq), len(k["rows"]))
print("instructions: ncu %d, nvdisasm %d" % (len(k["rows"]), len(seq)))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, collections.Counter()])
for i in range(n):
    r = k["rows"][i]; a = agg[seq[i]]
    a[0] += int(r[ci]); a[1] += int(r[si]); a[2] += int(r[ti]); a[3] += 1
    for c, h in stall_cols:
        try: a[4][h] += int(r[c])
        except ValueError: pass
ti_ = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("%-18s %7s %7s %6s %6s  top stalls" % ("function", "inst%", "stall%", "lanes", "sass"))
for f, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    st = ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(1, sum(a[4].values()))) for h, v in a[4].most_common(3))
    print("%-18s %6.1f%% %6.1f%% %6.1f %6d  %s" % (f, 100 * a[0] / ti_, 100 * a[1] / max(ts, 1), a[2] / max(a[0], 1), a[3], st))
