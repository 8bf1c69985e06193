#!/usr/bin/env python
"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm --print-line-info
to attribute executed instructions and stall samples to source lines.
usage: ncu_lines.py src.csv all.sass <mangled-kernel-substring> [file-substring]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
fsub = sys.argv[4] if len(sys.argv) > 4 else ""
# ---- nvdisasm: ordered list of (offset, line)
lines = open(sass).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # for inlined code keep the outermost "inlined at" location in our file if present
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
assert len(body) == len(seq), (len(body), len(seq))
agg = defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r, (off, loc, ins) in zip(body, seq):
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    key = loc if loc and (fsub in loc[0]) else ("other", 0)
    agg[key][0] += n
    agg[key][1] += s
    agg[key][2] += 1
    tot_i += n
    tot_s += s
print("total instr %d samples %d" % (tot_i, tot_s))
for k in sorted(agg, key=lambda k: (k[0], k[1])):
    a = agg[k]
    if a[0] * 200 > tot_i or a[1] * 200 > tot_s:
        print("%-28s line %4d  instr %6.2f%%  samples %6.2f%%  (%d sass)" % (k[0][-28:], k[1], 100.0 * a[0] / tot_i, 100.0 * a[1] / max(tot_s, 1), a[2]))
# optional region summary: pairs "name:lo-hi" after the file substring
if len(sys.argv) > 5:
    print("--- regions")
    for spec in sys.argv[5:]:
        name, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        ii = sum(a[0] for k, a in agg.items() if fsub in str(k[0]) and lo <= k[1] <= hi and k[0] != "other" and "ctc_loss_fast" in k[0])
        ss = sum(a[1] for k, a in agg.items() if lo <= k[1] <= hi and k[0] != "other" and "ctc_loss_fast" in k[0])
        print("%-16s instr %6.2f%%  samples %6.2f%%" % (name, 100.0 * ii / tot_i, 100.0 * ss / max(tot_s, 1)))
    ii = sum(a[0] for k, a in agg.items() if "ctc_loss_fast" not in str(k[0]))
    ss = sum(a[1] for k, a in agg.items() if "ctc_loss_fast" not in str(k[0]))
    print("%-16s instr %6.2f%%  samples %6.2f%%" % ("inlined/other", 100.0 * ii / tot_i, 100.0 * ss / max(tot_s, 1)))
