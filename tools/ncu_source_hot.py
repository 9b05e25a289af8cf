#!/usr/bin/env python3
"""Hot instructions of one kernel from `ncu -i rep --page source --csv`: stall-sample totals by reason and the top-N
instructions.  Usage: ncu_source_hot.py source.csv <kernel substring> [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
for b in blocks:
    if want not in b['name']:
        continue
    hdr, data = b['hdr'], b['rows']
    ix = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[ix['# Samples']]) for r in data)
    execd = sum(int(r[ix['Instructions Executed']]) for r in data)
    print(b['name'][:110])
    print("samples", tot, "static instrs", len(data), "warp-instrs executed", execd)
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in reasons}
    print("  by reason:", [(h[6:], v, "%.0f%%" % (100.0 * v / max(tot, 1))) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
    for i, r in sorted(enumerate(data), key=lambda t: -int(t[1][ix['# Samples']]))[:topn]:
        st = sorted(((h[6:], int(r[ix[h]])) for h in reasons), key=lambda kv: -kv[1])[:2]
        print("  %5d %6s smp %9s exec  %-58s %s xs=%s" % (i, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip()[:58], st, r[ix['L1 Wavefronts Shared Excessive']]))
