#!/usr/bin/env python
"""Per-kernel summary of the last forward + reverse step in an `ncu --metrics gpu__time_duration.sum` launch list."""
import csv, collections, sys
lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
idx = [i for i, r in enumerate(rows) if r['Kernel Name'].startswith('k_prepare')]
s = idx[-1]
agg = collections.OrderedDict(); tot = 0
for r in rows[s:]:
    t = float(r['Metric Value']) / 1e3
    a = agg.setdefault(r['Kernel Name'][:48] + " grid" + r['Grid Size'], [0, 0.0]); a[0] += 1; a[1] += t; tot += t
for k, v in agg.items():
    print("%-72s n=%3d  %9.1f us  %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("total us", tot)
