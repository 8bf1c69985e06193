#!/usr/bin/env python
"""ncu --csv launch list -> markdown table: ncu_launch_table.py in.csv "title" > out.md"""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 8]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
per = OrderedDict()
for r in rows[1:]:
    per.setdefault((r[ix['ID']], r[ix['Kernel Name']][:58], r[ix['Grid Size']] if 'Grid Size' in ix else ''), {})[r[ix['Metric Name']]] = (r[ix['Metric Value']], r[ix['Metric Unit']])
def val(m, n):
    v, u = m.get(n, ("0", ""))
    v = float(v.replace(",", ""))
    if u in ("ns", "nsecond"): v /= 1000
    if u == "msecond": v *= 1000
    if u == "byte": v /= 1e6
    if u == "Kbyte": v /= 1e3
    if u == "Gbyte": v *= 1e3
    return v
print("# " + sys.argv[2] + "\n")
print("| # | kernel | grid | us | tensor pipe active % | dram read MB | dram write MB |\n|---|---|---|---|---|---|---|")
tot = 0
for (i, k, g), m in per.items():
    t = val(m, 'gpu__time_duration.sum'); tot += t
    print("| %s | %s | %s | %.1f | %.1f | %.2f | %.2f |" % (i, k, g, t, val(m, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), val(m, 'dram__bytes_read.sum'), val(m, 'dram__bytes_write.sum')))
print("\ntotal %.1f us (serialised under ncu, cold cache: compare shares)" % tot)
