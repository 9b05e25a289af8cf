#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
Usage: launch_summary.py launches.csv [last_n_launches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
recs = [(r[ki], r[gi], float(r[vi].replace(',', ''))) for r in rows[h + 1:] if len(r) > vi and r[vi].replace(',', '').replace('.', '').isdigit()]
if len(sys.argv) > 2:
    recs = recs[-int(sys.argv[2]):]
tot = sum(t for _, _, t in recs)
agg = collections.OrderedDict()
for name, grid, t in recs:
    m = re.search(r'kernel_entry<(?:hf::)?([A-Za-z0-9_]+(?:<[^>]*>)?)', name)
    k = m.group(1) if m else name[:40]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += t
print("%d launches, %.3f ms total (serialised, cold cache)" % (len(recs), tot / 1e6))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %4d launches %9.3f ms %5.1f %%" % (k, n, t / 1e6, 100 * t / tot))
