#!/usr/bin/env python
"""ncu --csv launch list (gpu__time_duration.sum) -> per-kernel totals: ncu_by_kernel.py in.csv "title" [skip_launches] > out.md"""
import csv, sys, re
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
agg = OrderedDict(); tot = 0.0; n = 0
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum': continue
    n += 1
    if n <= skip: continue
    v = float(r[ix['Metric Value']].replace(",", "")); u = r[ix['Metric Unit']]
    us = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[ix['Kernel Name']])[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
print("# " + sys.argv[2] + "\n")
print("| kernel | launches | total us | share % |\n|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f |" % (k, c, t, 100 * t / tot))
print("\ntotal %.1f us over %d launches (serialised under ncu, cold cache: compare shares)" % (tot, n - skip))
