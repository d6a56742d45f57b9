#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass` output.  usage: ncu_lines.py file.csv [min_pct]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
fname = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[2] != '-': continue      # sass rows carry an address; '-' rows are the per-line totals
    ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
    k = (fname, int(r[0]))
    a = agg.setdefault(k, [0, 0, r[1]])
    a[0] += int(r[ii] or 0); a[1] += int(r[si] or 0)
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
print("total inst %d samples %d" % (tot, tots))
for (f, l), a in agg.items():
    if 100.0 * a[0] / tot >= minp or 100.0 * a[1] / tots >= minp:
        print("%-14s %5d %6.2f%% inst %6.2f%% smp  %s" % (f, l, 100.0 * a[0] / tot, 100.0 * a[1] / tots, a[2].strip()[:100]))
