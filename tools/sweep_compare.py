#!/usr/bin/env python
"""Side-by-side medians of several tools/sweep.py --full outputs (forward / backward us per grid cell)."""
import json
import sys

files = sys.argv[1:]
data = [json.load(open(f)) for f in files]
grids = [{(g['nnz'], g['C'], g['skew']): g for g in d['grid']} for d in data]
print("%8s %3s %-8s | " % ("nnz", "C", "skew") + " | ".join("%-17s" % f.split('/')[-1].replace('.json', '')[-17:] for f in files))
for k in sorted(grids[0]):
    cells = []
    for g in grids:
        x = g.get(k)
        cells.append("%7.1f %7.1f  " % (x['fwd_us_median'], x['bwd_us_median']) if x else " " * 17)
    print("%8d %3d %-8s | " % k + " | ".join(cells))
rows = [{r['case']: r for r in d['rows'] if not r['case'].startswith('cfg5')} for d in data]
for k in rows[0]:
    print("%-28s | " % k[:28] + " | ".join("%7.1f %7.1f  " % (r[k]['fwd_us'], r[k]['bwd_us']) if k in r else " " * 17 for r in rows))
