#!/usr/bin/env python
"""Top source lines of a kernel by samples / executed instructions, from
   ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > src.csv
   python scripts/ncu_by_line.py src.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = None; hdr = None; agg = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue
    isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
    try:
        key = (fname, int(r[0]))
        s, e = int(r[isamp] or 0), int(r[iex] or 0)
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += s; a[1] += e
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print("lines %d  samples %d  executed %.3e" % (len(agg), ts, te))
for (f, l), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% smp %5.2f%% ins  %s:%d  %s" % (100 * s / ts, 100 * e / te, f, l, src.strip()[:110]))
