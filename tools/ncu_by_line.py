#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <cubin> <ncu-kernel-name-substring> [top] [cubin-section-substring]
Joins `ncu --page source --print-source sass --csv` (per-instruction counters) with `nvdisasm -g` line info
of the same cubin (instruction order is identical), and prints the hottest source lines by executed
warp-instructions and by stall samples.  Needs the build's -lineinfo.
"""
import csv, re, subprocess, sys, collections, io

rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sect = sys.argv[5] if len(sys.argv) > 5 else "env_kernel"   # substring of the cubin .text section (mangled name)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# split per kernel
kernels = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, rows=[]); kernels.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
k = [x for x in kernels if kern in x["name"]][0]
hdr = k["hdr"]; ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed")
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# find the function section for our kernel
lines = dis.split("\n")
infn = False; cur_line = ("?", 0); seq = []
for ln in lines:
    if ln.startswith("//--------------------- .text."):
        infn = sect in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        seq.append(cur_line)
n = min(len(seq), len(k["rows"]))
print("instructions: ncu %d, nvdisasm %d" % (len(k["rows"]), len(seq)))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for i in range(n):
    r = k["rows"][i]
    a = agg[seq[i]]
    a[0] += int(r[ci]); a[1] += int(r[si]); a[2] += int(r[ti]); a[3] += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
src_cache = {}
def src(f, l):
    import os
    for d in ("ur3e_b200/csrc",):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().split("\n")
            return src_cache[p][l - 1].strip()[:110] if l - 1 < len(src_cache[p]) else ""
    return ""
print("\n== by executed warp-instructions")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  lanes %4.1f  sass %4d  %s:%d  %s" % (100 * a[0] / tot_i, 100 * a[1] / max(tot_s, 1), a[2] / max(a[0], 1), a[3], f, l, src(f, l)))
print("\n== by stall samples")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100 * a[1] / max(tot_s, 1), 100 * a[0] / tot_i, f, l, src(f, l)))
