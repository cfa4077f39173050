#!/usr/bin/env python
"""Aggregate the warp-stall samples of an .ncu-rep (captured with --import-source on, built with -lineinfo) per CUDA source line.
usage: ncu_hot_lines.py report.ncu-rep [min_pct]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
per = collections.OrderedDict()
stalls = collections.defaultdict(lambda: collections.Counter())
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iS = hdr.index("# Samples")
        st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].strip():
        cur = (r[0], r[1])
        per.setdefault(cur, 0)
    if r[2].strip() and cur is not None and r[iS].isdigit():
        per[cur] += int(r[iS])
        for i, h in st_cols:
            if r[i].isdigit():
                stalls[cur][h] += int(r[i])
tot = sum(per.values())
print("total samples", tot)
for k, v in per.items():
    if tot and 100.0 * v / tot >= minpct:
        top = ", ".join("%s %d" % (h[6:], c) for h, c in stalls[k].most_common(3))
        print("%5s %6d %5.1f%%  %-110s | %s" % (k[0], v, 100.0 * v / tot, k[1].strip()[:110], top))
